/*
 * bhs.h -- C ABI of libbhs, the B200 (sm_100a) implementation of the biem_helmholtz_sphere hot path.
 *
 * The reference (ultrasphere-dev/biem-helmholtz-sphere v1.2.0) has NO FFI / plugin interface: its
 * boundary is the Python API `biem(...)` -> `BIEMResultCalculator.uscat(...)`
 * (src/biem_helmholtz_sphere/_biem.py:453, :223, :822).  Each entry point below replaces the Python
 * call sites of that file that it cites; `biem_helmholtz_sphere_b200/_biem.py` is the host-side
 * mirror of the reference API that binds them through ctypes.
 *
 * Conventions
 *   - every pointer marked `d_` is a DEVICE pointer; complex128 = interleaved (re, im) doubles;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising; calls are re-entrant, there is no hidden global state except the plan cache
 *     objects the caller owns;
 *   - return value: 0 = ok, <0 = invalid argument (BHS_ERR_*), >0 = cudaError_t of a failed launch;
 *   - coordinate types are the chain trees 'a' (d=2), 'ba' (d=3), 'bba' (d=4), 'bbba' (d=5), ... up to d=8: passed as `d`;
 *   - wavenumbers are real arrays `d_k`; an optional `d_k_im` (NULL = real wavenumbers) adds imaginary parts
 *     (absorbing media, Im k != 0); h_n^{(1)} is then always computed directly (never as j + i y, which cancels);
 *   - matrices are ROW-major (C order), exactly the reference's [..., B, harm, B', harm'] layout.
 */
#ifndef BHS_H
#define BHS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BHS_OK 0
#define BHS_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, ...)            */
#define BHS_ERR_UNSUPPORTED (-2) /* valid request outside the implemented range                */
#define BHS_ERR_ALLOC (-3)       /* host or device allocation failed while building a plan     */

#define BHS_KIND_J 0  /* regular   j_n^{(d)}                                                    */
#define BHS_KIND_Y 1  /* irregular y_n^{(d)}                                                    */
#define BHS_KIND_H1 2 /* Hankel    h_n^{(d)} = j + i y                                          */

#define BHS_FLAG_PER_BALL 1 /* uscat: keep the ball axis   (_biem.py:964)                        */
#define BHS_FLAG_FAR_FIELD 2 /* uscat: far-field pattern   (_biem.py:930-959)                    */
#define BHS_FLAG_INNER 4     /* uscat: kind == "inner" NaN mask (_biem.py:973-974)               */

typedef struct bhs_plan bhs_plan_t;

/* library / device ------------------------------------------------------------------------- */
int bhs_version(void);
/* Number of SMs of the current device (used by host code to size persistent grids). */
int bhs_device_sm_count(int *out);

/* plans: k-independent tables for one (d, n_end) --------------------------------------------
 * Replaces ush.index_array_harmonics / flatten_harmonics (_biem.py:651,720,743,889,917),
 * the quadrature inside ush.expand (_biem.py:627) and the coupling (Gaunt-type) coefficients
 * inside ush.harmonics_translation_coef (_biem.py:697).  Built on the host in extended precision,
 * uploaded once; the coupling table is stored tile-wise for TMA staging by bhs_assemble. */
int bhs_plan_create(int d, int n_end, bhs_plan_t **out);
/* Coordinate trees other than the chains.  Every tree of one dimension spans the same harmonic space (degree < n_end on
 * S^{d-1}); what a tree changes in the reference is (i) the labelling / orientation of the cartesian axes -- trees with b'
 * nodes are chains in a permuted frame, handled by the host code -- and (ii) where ush.expand samples the sphere
 * (_biem.py:627).  BHS_TREE_HOPF (d = 4, the reference's 'caa': type-c root over two type-a circles) builds a plan whose
 * right-hand-side quadrature is the Hopf product rule (n_end Gauss-Legendre nodes in cos 2 theta_0, 2 n_end equispaced nodes
 * per circle; pinned by jascome/jascome_output.csv:2-6); it serves bhs_rhs_expand / bhs_plan_quadrature only. */
#define BHS_TREE_CHAIN 0
#define BHS_TREE_HOPF 1
int bhs_plan_create_tree(int d, int n_end, int tree, bhs_plan_t **out);
void bhs_plan_destroy(bhs_plan_t *plan);
int bhs_plan_harm(const bhs_plan_t *plan);       /* H  = number of harmonics of degree < n_end   */
int bhs_plan_harm2(const bhs_plan_t *plan);      /* H2 = number of harmonics of degree < 2n_end-1 */
int bhs_plan_quad_points(const bhs_plan_t *plan); /* Q  = nodes of the RHS product rule           */
/* Copies the flattened index table [H, d-1] (columns n_0, n_1, ..., signed m) to HOST memory. */
int bhs_plan_index_table(const bhs_plan_t *plan, int32_t *h_out);
/* Copies the unit quadrature directions [d, Q] and weights [Q] to HOST memory. */
int bhs_plan_quadrature(const bhs_plan_t *plan, double *h_dirs, double *h_weights);
/* Number of (entry, term) pairs in the coupling table and its device footprint in bytes. */
int bhs_plan_coupling_stats(const bhs_plan_t *plan, int64_t *nterms, int64_t *bytes);

/* K1: hyperspherical Bessel / Hankel functions ------------------------------------------------
 * out[i, n] (n = 0..n_max) = z_n^{(d)}(x_i) = sqrt(pi/2) Z_{n+d/2-1}(x_i) / x_i^{d/2-1}, or d/dx.
 * Replaces ultrasphere.shn1 / potential_coef / harmonics_regular_singular_component
 * (_biem.py:439,447,654-685,723-741,750-787,896-914).  Real positive arguments.
 * d_out: complex128 [nx, n_max+1] (imaginary part zero for J and Y). */
int bhs_bessel(int d, int kind, int derivative, int n_max, const double *d_x, int64_t nx,
               double *d_out, void *stream);

/* Complex argument x = x_re + i x_im != 0 (complex wavenumbers): kind J or H1 only, same layout.  j by Miller's
 * algorithm normalised on closed forms (spherical) / on e^{-+ix} = J_0 + 2 sum (-+i)^k J_k (cylindrical); h^{(1)} by
 * upward recurrence from orders 0, 1 (spherical: closed forms; cylindrical: Hankel asymptotics for |x| >= 18,
 * K_nu(-ix) by Steed's CF2 for Im x > 3, J + iY otherwise). */
int bhs_bessel_z(int d, int kind, int derivative, int n_max, const double *d_x_re,
                 const double *d_x_im, int64_t nx, double *d_out, void *stream);

/* K2: orthonormal harmonics in ultrasphere (Phase(0)) order ----------------------------------
 * d_xyz: [d, npts] cartesian points (need not be normalised); d_out: complex128 [npts, H(n_end2)].
 * n_end2 may be n_end or 2*n_end-1 of the plan.  Replaces ush.harmonics (_biem.py:922). */
int bhs_harmonics(const bhs_plan_t *plan, int use_double_band, const double *d_xyz, int64_t npts,
                  double *d_out, void *stream);

/* K3: right-hand side  f_hat[s, b, h] = sum_q w_q g[s, q, b] conj Y_h(y_q) ----------------------
 * Replaces ush.expand(..., n=n_end) (_biem.py:627-639).
 *   d_g != NULL : boundary data sampled by the caller, complex128 [nsys, Q, B]
 *                 (= -alpha*uin - beta*d_n uin at rho_b*y_q + c_b, _biem.py:611-624);
 *   d_g == NULL : fused plane wave exp(i k_in dir.x) (_biem.py:375-386) with per-ball complex
 *                 alpha/beta [B] (NULL = 1 / 0); d_k_in real [nsys] wavenumber of the INCIDENT wave.
 * d_centers [B, d], d_radii [B], d_dir [d] (unit), d_out complex128 [nsys, B, H]. */
int bhs_rhs_expand(const bhs_plan_t *plan, int B, int nsys, const double *d_g,
                   const double *d_centers, const double *d_radii, const double *d_k_in,
                   const double *d_k_in_im, const double *d_dir, const double *d_alpha,
                   const double *d_beta, double *d_out, void *stream);

/* K4: system assembly --------------------------------------------------------------------------
 * A[s][(b,h),(b',h')] = SD_{n'}(rho_b') * ( b==b' ? delta (alpha h_n + beta k h_n')(k rho_b)
 *                                                 : (S|R)_{h',h}(c_b - c_b') (alpha j_n + beta k j_n')(k rho_b) )
 * Replaces _biem.py:692-792 (harmonics_translation_coef, potential_coef, create_diagonal, where).
 * d_k, d_eta real [nsys]; alpha/beta complex [B] or NULL; d_A complex128 [nsys][N, ld], N = B*H,
 * ld >= N (elements); sys_stride in complex elements.  d_work: bhs_assemble_workspace() bytes. */
int64_t bhs_assemble_workspace(const bhs_plan_t *plan, int B, int nsys);
int bhs_assemble(const bhs_plan_t *plan, int B, int nsys, const double *d_centers,
                 const double *d_radii, const double *d_k, const double *d_k_im, const double *d_eta,
                 const double *d_alpha, const double *d_beta, double *d_A, int64_t ld,
                 int64_t sys_stride, void *d_work, void *stream);

/* Block rows b in [b_lo, b_hi) only, into a strip [(b_hi - b_lo) * H, ld] per system whose first row is (b_lo, h = 0):
 * the unit of the multi-GPU assembly block-row sharding.  Same workspace as bhs_assemble. */
int bhs_assemble_rows(const bhs_plan_t *plan, int B, int nsys, const double *d_centers,
                      const double *d_radii, const double *d_k, const double *d_k_im,
                      const double *d_eta, const double *d_alpha, const double *d_beta, int b_lo, int b_hi,
                      double *d_A, int64_t ld, int64_t sys_stride, void *d_work, void *stream);

/* single-sphere shortcut diag[s, b, h] = SD_n (alpha h_n + beta k h_n')  (_biem.py:648-691);
 * d_work: bhs_diag_coef_workspace() bytes (the library never allocates device memory inside an entry point) */
int64_t bhs_diag_coef_workspace(const bhs_plan_t *plan, int B, int nsys);
int bhs_diag_coef(const bhs_plan_t *plan, int B, int nsys, const double *d_radii, const double *d_k,
                  const double *d_k_im, const double *d_eta, const double *d_alpha, const double *d_beta,
                  double *d_out, void *d_work, void *stream);

/* K5: dense complex128 solve, row-major, blocked LU with tournament partial pivoting, trailing
 * update on FP64 tensor cores (DMMA).  Replaces batch_tensorsolve.btensorsolve -> zgesv
 * (_biem.py:797).  A is overwritten by its LU factors, rhs [N, nrhs] (row-major, ld = nrhs) by the
 * solution.  d_ipiv int32 [N] (0-based row swapped with row i at step i); d_info int32 [1]
 * (0 ok, i+1 = zero pivot at step i).  d_work: bhs_zgesv_workspace(N, nrhs) bytes. */
int64_t bhs_zgesv_workspace(int64_t N, int nrhs);
int bhs_zgesv(int64_t N, int nrhs, double *d_A, int64_t ld, double *d_rhs, int32_t *d_ipiv,
              int32_t *d_info, void *d_work, void *stream);
/* The same for `nbatch` systems of one size in lock step (one launch per step for the whole group: what a wavenumber
 * sweep issues).  A [nbatch][N, ld] with strideA complex elements between systems, rhs [nbatch][N, nrhs] stride_rhs apart,
 * d_ipiv int32 [nbatch][N], d_info int32 [nbatch]. */
int64_t bhs_zgesv_batched_workspace(int64_t N, int nrhs, int nbatch);
int bhs_zgesv_batched(int64_t N, int nrhs, int nbatch, double *d_A, int64_t ld, int64_t strideA,
                      double *d_rhs, int64_t stride_rhs, int32_t *d_ipiv, int32_t *d_info, void *d_work,
                      void *stream);
/* The pieces, exposed for tests and for re-solving with stored factors. */
int bhs_zgetrf(int64_t N, double *d_A, int64_t ld, int32_t *d_ipiv, int32_t *d_info, void *d_work,
               void *stream);
int bhs_zgetrs(int64_t N, int nrhs, const double *d_LU, int64_t ld, const int32_t *d_ipiv,
               double *d_rhs, void *d_work, void *stream);
/* C[M,N] -= A[M,K] * B[K,N] (all row-major complex128) on DMMA: the LU trailing update, exposed
 * for the roofline measurement.  K must be a multiple of 8. */
int64_t bhs_zgemm_workspace(int64_t M, int64_t N, int64_t K);
int bhs_zgemm_sub(int64_t M, int64_t N, int64_t K, const double *d_A, int64_t lda,
                  const double *d_B, int64_t ldb, double *d_C, int64_t ldc, void *d_work,
                  void *stream);

/* K6: scattered field  u_s(x) = sum_b sum_h density[b,h] SD_n(rho_b) h_n(k|x-c_b|) Y_h((x-c_b)^) ----
 * Replaces biem_u (_biem.py:822-977).  d_x [d, P]; d_centers [B, d]; d_density complex128 [B, H];
 * d_out complex128 [P] (or [P, B] with BHS_FLAG_PER_BALL).  NaN inside any ball (outside for
 * BHS_FLAG_INNER) unless BHS_FLAG_FAR_FIELD. d_work: bhs_uscat_workspace() bytes. */
int64_t bhs_uscat_workspace(const bhs_plan_t *plan, int B);
int bhs_uscat(const bhs_plan_t *plan, int B, const double *d_centers, const double *d_radii,
              double k, double k_im, double eta, const double *d_density, const double *d_x, int64_t P,
              int flags, double *d_out, void *d_work, void *stream);

/* measurement helpers (used by bench.py) ------------------------------------------------------- */
/* Kernel launches issued by the library since the last reset (host-side count; graph replays of a
 * captured sequence are not seen here -- bench.py multiplies the eager count). */
int64_t bhs_launch_count(int reset);
/* Per-category device timing: bhs_profile(1) clears and enables CUDA-event brackets around the launch
 * groups below (not usable during graph capture), bhs_profile(0) disables.  bhs_profile_read
 * synchronises and returns the summed milliseconds, the summed algorithmic work (flops or bytes, see
 * DESIGN.md) and the number of brackets of one category. */
#define BHS_PROF_LU_GEMM 0      /* zgemm_sub_kernel, trailing updates of the outer blocks, K = 128
                                   (work = 8 M N K flops; 97.7 % of the flops of a factorisation)  */
#define BHS_PROF_LU_PANEL 1     /* tournament pivoting + swaps + diagonal LU + L21                */
#define BHS_PROF_LU_TRSM 2      /* U12 triangular solves                                          */
#define BHS_PROF_LU_PACK 3      /* operand packing for the DMMA kernel                            */
#define BHS_PROF_LU_RHS 4       /* forward / backward substitution of the right-hand sides        */
#define BHS_PROF_ASM_MAIN 5     /* assemble_kernel (work = 16 N^2 bytes)                          */
#define BHS_PROF_ASM_PRE 6      /* radial tables, pair harmonics, factors                         */
#define BHS_PROF_USCAT 7        /* field kernel (work = 8 P B H flops)                            */
#define BHS_PROF_RHS_EXPAND 8   /* right-hand-side expansion                                      */
#define BHS_PROF_LU_GEMM_IN 9   /* zgemm_sub_kernel, K = 32 / 64 updates inside the panel recursion  */
#define BHS_PROF_NCAT 10
int bhs_profile(int enable);
int bhs_profile_read(int category, double *ms, double *work, int64_t *count);

/* Runs a register-resident DFMA loop / DMMA loop on every SM; returns achieved TFLOP/s. shape:
 * 0 = DFMA, 1 = mma.m8n8k4.f64, 2 = m16n8k4, 3 = m16n8k8, 4 = m16n8k16. */
int bhs_fp64_peak(int shape, int iters, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* BHS_H */
