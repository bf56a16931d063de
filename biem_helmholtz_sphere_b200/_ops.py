"""Thin torch wrappers over the C ABI (include/bhs.h): allocate outputs / workspaces with torch's caching
allocator, pass raw device pointers and the current stream.  Everything here runs on the current CUDA
device; there is no CPU path."""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, get_plan, load, ptr, stream_ptr

F64 = torch.float64
C128 = torch.complex128


def _dev():
    if not torch.cuda.is_available():
        raise _lib.BhsError("biem_helmholtz_sphere_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _f64(x) -> torch.Tensor:
    return torch.as_tensor(x, dtype=F64, device=_dev()).contiguous()


def _c128(x) -> torch.Tensor:
    return torch.as_tensor(x, dtype=C128, device=_dev()).contiguous()


def _work(nbytes: int) -> torch.Tensor:
    if nbytes < 0:
        check(int(nbytes), "workspace query")
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=_dev())


def bessel(d: int, kind: int, n_max: int, x, derivative: bool = False) -> torch.Tensor:
    """z_n^{(d)}(x) for n = 0..n_max -> complex128 [*x.shape, n_max+1]   (bhs_bessel)."""
    xt = _f64(x)
    out = torch.empty(xt.shape + (n_max + 1,), dtype=C128, device=xt.device)
    check(load().bhs_bessel(d, kind, int(derivative), n_max, ptr(xt), xt.numel(), ptr(out), stream_ptr()), "bhs_bessel")
    return out


def bessel_z(d: int, kind: int, n_max: int, z, derivative: bool = False) -> torch.Tensor:
    """z_n^{(d)}(z) for complex arguments z, n = 0..n_max (kind J or H1) -> complex128 [*z.shape, n_max+1]   (bhs_bessel_z)."""
    zt = torch.as_tensor(z, device=_dev()).to(C128)
    zr, zi = zt.real.contiguous(), zt.imag.contiguous()
    out = torch.empty(zt.shape + (n_max + 1,), dtype=C128, device=zt.device)
    check(load().bhs_bessel_z(d, kind, int(derivative), n_max, ptr(zr), ptr(zi), zt.numel(), ptr(out), stream_ptr()),
          "bhs_bessel_z")
    return out


def harmonics(d: int, n_end: int, xyz, double_band: bool = False) -> torch.Tensor:
    """Y_h at the directions of xyz [d, ...] -> complex128 [..., H]   (bhs_harmonics)."""
    plan = get_plan(d, n_end)
    xt = _f64(xyz)
    shape = xt.shape[1:]
    xt = xt.reshape(d, -1).contiguous()
    Hb = plan.H2 if double_band else plan.H
    out = torch.empty((xt.shape[1], Hb), dtype=C128, device=xt.device)
    check(load().bhs_harmonics(plan.handle, int(double_band), ptr(xt), xt.shape[1], ptr(out), stream_ptr()), "bhs_harmonics")
    return out.reshape(shape + (Hb,))


def rhs_expand(d: int, n_end: int, *, g=None, centers=None, radii=None, k_in=None, direction=None, alpha=None,
               beta=None, B: int | None = None, k_in_im=None, tree: int = 0) -> torch.Tensor:
    """f_hat [nsys, B, H]; either sampled boundary data g [nsys, Q, B] or a fused plane wave exp(i (k_in + i k_in_im) d.x).
    ``tree``: whose quadrature samples the sphere (_lib.TREE_CHAIN, or _lib.TREE_HOPF for the 'caa' tree)."""
    plan = get_plan(d, n_end, tree)
    if g is not None:
        gt = _c128(g)
        nsys, Q, B = gt.shape
        if Q != plan.Q:
            raise ValueError(f"g must have {plan.Q} quadrature rows, got {Q}")
        out = torch.empty((nsys, B, plan.H), dtype=C128, device=gt.device)
        check(load().bhs_rhs_expand(plan.handle, B, nsys, ptr(gt), None, None, None, None, None, None, None, ptr(out),
                                    stream_ptr()), "bhs_rhs_expand")
        return out
    cen, rad, kk, dr = _f64(centers), _f64(radii), _f64(k_in).reshape(-1), _f64(direction)
    B = rad.shape[0]
    al = None if alpha is None else _c128(alpha)
    be = None if beta is None else _c128(beta)
    kim = None if k_in_im is None else _f64(k_in_im).reshape(-1)
    out = torch.empty((kk.numel(), B, plan.H), dtype=C128, device=cen.device)
    check(load().bhs_rhs_expand(plan.handle, B, kk.numel(), None, ptr(cen), ptr(rad), ptr(kk), ptr(kim), ptr(dr), ptr(al),
                                ptr(be), ptr(out), stream_ptr()), "bhs_rhs_expand")
    return out


def assemble(d: int, n_end: int, centers, radii, k, eta=None, alpha=None, beta=None, out=None, work=None,
             k_im=None, rows: tuple[int, int] | None = None) -> torch.Tensor:
    """A [nsys, N, N] row-major   (bhs_assemble); wavenumbers k + i k_im (k_im None = real).

    ``rows=(b_lo, b_hi)``: only the block rows of the row balls b_lo <= b < b_hi, as a strip [nsys, (b_hi-b_lo) H, N]
    (bhs_assemble_rows: the unit of the multi-GPU block-row sharding)."""
    plan = get_plan(d, n_end)
    cen, rad, kk = _f64(centers), _f64(radii), _f64(k).reshape(-1)
    B = rad.shape[0]
    nsys = kk.numel()
    N = B * plan.H
    et = None if eta is None else _f64(eta).reshape(-1)
    al = None if alpha is None else _c128(alpha)
    be = None if beta is None else _c128(beta)
    b_lo, b_hi = (0, B) if rows is None else rows
    nrow = (b_hi - b_lo) * plan.H
    if out is None:
        out = torch.empty((nsys, nrow, N), dtype=C128, device=cen.device)
    if work is None:
        work = _work(load().bhs_assemble_workspace(plan.handle, B, nsys))
    kim = None if k_im is None else _f64(k_im).reshape(-1)
    if rows is None:
        check(load().bhs_assemble(plan.handle, B, nsys, ptr(cen), ptr(rad), ptr(kk), ptr(kim), ptr(et), ptr(al), ptr(be),
                                  ptr(out), N, N * N, ptr(work), stream_ptr()), "bhs_assemble")
    else:
        check(load().bhs_assemble_rows(plan.handle, B, nsys, ptr(cen), ptr(rad), ptr(kk), ptr(kim), ptr(et), ptr(al),
                                       ptr(be), b_lo, b_hi, ptr(out), N, nrow * N, ptr(work), stream_ptr()),
              "bhs_assemble_rows")
    return out


def diag_coef(d: int, n_end: int, radii, k, eta=None, alpha=None, beta=None, k_im=None) -> torch.Tensor:
    plan = get_plan(d, n_end)
    rad, kk = _f64(radii), _f64(k).reshape(-1)
    B = rad.shape[0]
    et = None if eta is None else _f64(eta).reshape(-1)
    al = None if alpha is None else _c128(alpha)
    be = None if beta is None else _c128(beta)
    out = torch.empty((kk.numel(), B, plan.H), dtype=C128, device=rad.device)
    kim = None if k_im is None else _f64(k_im).reshape(-1)
    work = _work(load().bhs_diag_coef_workspace(plan.handle, B, kk.numel()))
    check(load().bhs_diag_coef(plan.handle, B, kk.numel(), ptr(rad), ptr(kk), ptr(kim), ptr(et), ptr(al), ptr(be),
                               ptr(out), ptr(work), stream_ptr()), "bhs_diag_coef")
    return out


class SolveBuffers:
    """Reusable ipiv / info / workspace of bhs_zgesv (or bhs_zgesv_batched with nbatch systems) for one system size."""

    def __init__(self, N: int, nrhs: int = 1, nbatch: int = 1):
        dev = _dev()
        self.N = N
        self.nbatch = nbatch
        self.ipiv = torch.empty((nbatch * N,), dtype=torch.int32, device=dev)
        self.info = torch.zeros((nbatch,), dtype=torch.int32, device=dev)
        self.work = _work(load().bhs_zgesv_batched_workspace(N, nrhs, nbatch))


def zgesv_(A: torch.Tensor, rhs: torch.Tensor, bufs: SolveBuffers | None = None):
    """In-place solve: A [N, N] (row-major, overwritten by LU), rhs [N] or [N, nrhs] (overwritten by x)."""
    N = A.shape[0]
    assert A.dtype == C128 and A.is_cuda and A.stride(1) == 1
    assert rhs.dtype == C128 and rhs.is_contiguous()
    nrhs = 1 if rhs.dim() == 1 else rhs.shape[1]
    if bufs is None:
        bufs = SolveBuffers(N, nrhs)
    check(load().bhs_zgesv(N, nrhs, ptr(A), A.stride(0), ptr(rhs), ptr(bufs.ipiv), ptr(bufs.info), ptr(bufs.work),
                           stream_ptr()), "bhs_zgesv")
    return rhs, bufs


def zgesv_batched_(A: torch.Tensor, rhs: torch.Tensor, bufs: SolveBuffers | None = None):
    """In-place solve of S independent systems in lock step: A [S, N, N] (overwritten by the factors), rhs [S, N] or
    [S, N, nrhs] (overwritten by the solutions)   (bhs_zgesv_batched)."""
    S, N = A.shape[0], A.shape[1]
    assert A.dtype == C128 and A.is_cuda and A.stride(2) == 1 and rhs.dtype == C128 and rhs.is_contiguous()
    nrhs = 1 if rhs.dim() == 2 else rhs.shape[2]
    if bufs is None:
        bufs = SolveBuffers(N, nrhs, S)
    assert bufs.nbatch >= S
    check(load().bhs_zgesv_batched(N, nrhs, S, ptr(A), A.stride(1), A.stride(0), ptr(rhs), rhs.stride(0), ptr(bufs.ipiv),
                                   ptr(bufs.info), ptr(bufs.work), stream_ptr()), "bhs_zgesv_batched")
    return rhs, bufs


def zgetrf_(A: torch.Tensor, bufs: SolveBuffers | None = None):
    N = A.shape[0]
    if bufs is None:
        bufs = SolveBuffers(N, 1)
    check(load().bhs_zgetrf(N, ptr(A), A.stride(0), ptr(bufs.ipiv), ptr(bufs.info), ptr(bufs.work), stream_ptr()),
          "bhs_zgetrf")
    return bufs


def zgetrs_(LU: torch.Tensor, bufs: SolveBuffers, rhs: torch.Tensor):
    N = LU.shape[0]
    nrhs = 1 if rhs.dim() == 1 else rhs.shape[1]
    check(load().bhs_zgetrs(N, nrhs, ptr(LU), LU.stride(0), ptr(bufs.ipiv), ptr(rhs), ptr(bufs.work), stream_ptr()),
          "bhs_zgetrs")
    return rhs


def zgemm_sub_(Cm: torch.Tensor, A: torch.Tensor, Bm: torch.Tensor, work=None):
    """C -= A @ B on the DMMA kernel (row-major complex128)."""
    M, K = A.shape
    K2, N = Bm.shape
    assert K == K2 and Cm.shape == (M, N)
    if work is None:
        work = _work(load().bhs_zgemm_workspace(M, N, K))
    check(load().bhs_zgemm_sub(M, N, K, ptr(A), A.stride(0), ptr(Bm), Bm.stride(0), ptr(Cm), Cm.stride(0), ptr(work),
                               stream_ptr()), "bhs_zgemm_sub")
    return Cm


def uscat(d: int, n_end: int, centers, radii, k: float | complex, eta: float, density, x, *, far_field=False,
          per_ball=False, inner=False, work=None) -> torch.Tensor:
    """u_s at x [d, P] -> [P] (or [P, B])   (bhs_uscat); k may be complex (3-D only)."""
    plan = get_plan(d, n_end)
    cen, rad, den, xt = _f64(centers), _f64(radii), _c128(density), _f64(x)
    B = rad.shape[0]
    P = xt.shape[1]
    flags = (_lib.FLAG_PER_BALL if per_ball else 0) | (_lib.FLAG_FAR_FIELD if far_field else 0) | (
        _lib.FLAG_INNER if inner else 0)
    out = torch.empty((P, B) if per_ball else (P,), dtype=C128, device=xt.device)
    if work is None:
        work = _work(load().bhs_uscat_workspace(plan.handle, B))
    kc = complex(k)
    check(load().bhs_uscat(plan.handle, B, ptr(cen), ptr(rad), kc.real, kc.imag, float(eta), ptr(den), ptr(xt), P, flags,
                           ptr(out), ptr(work), stream_ptr()), "bhs_uscat")
    return out


def fp64_peak(shape: int, iters: int = 4096) -> float:
    v = C.c_double()
    check(load().bhs_fp64_peak(shape, iters, C.byref(v)), "bhs_fp64_peak")
    return v.value


def launch_count(reset: bool = False) -> int:
    """Kernel launches issued by libbhs since the last reset (bhs_launch_count)."""
    return int(load().bhs_launch_count(int(reset)))


def profile(enable: bool) -> None:
    """Enable / disable (and clear) the per-category CUDA-event profiler (bhs_profile)."""
    check(load().bhs_profile(int(enable)), "bhs_profile")


def profile_read() -> dict:
    """{category: {"ms", "work", "count"}} accumulated since profile(True)   (bhs_profile_read)."""
    out = {}
    for i, name in enumerate(_lib.PROF_CATEGORIES):
        ms, wk, n = C.c_double(), C.c_double(), C.c_int64()
        check(load().bhs_profile_read(i, C.byref(ms), C.byref(wk), C.byref(n)), "bhs_profile_read")
        out[name] = {"ms": ms.value, "work": wk.value, "count": n.value}
    return out
