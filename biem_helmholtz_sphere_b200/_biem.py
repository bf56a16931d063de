"""Host-side mirror of the reference's public API for the BIEM hot path.

Same names, signatures, argument meaning and error behaviour as
``/root/reference/src/biem_helmholtz_sphere/_biem.py`` (``biem`` :453, ``BIEMResultCalculator`` :196,
``biem_u`` :822, ``plane_wave`` :329, ``point_source`` :391, ``max_memory`` :23, ``max_n_end`` :52), but every
numerical step is a call into libbhs.so (hand-written sm_100a kernels, include/bhs.h).  PyTorch is used for
device memory, streams and CUDA graphs only.  There is no CPU fallback: without a CUDA device or without the
built library the calls raise.

Inputs may be NumPy arrays or torch tensors (any device); results come back in the caller's array namespace
(NumPy in -> NumPy out), computed in float64 / complex128 on the current CUDA device.

Deliberate differences from the reference (documented in DESIGN.md):
* float32 inputs are up-cast; outputs are always complex128 (north_star: complex128 only);
* complex wavenumbers (Im k != 0) take their own kernels: h_n^{(1)} by complex upward recurrence from orders 0, 1 (never
  as j + i y), j_n by Miller's algorithm -- oracle-pinned against scipy's complex-argument Bessel functions;
* leading batch axes of ``k`` WORK together with ``uin`` (the reference raises there); the
  semantics are "identical to a loop of scalar-k calls";
* extra keyword ``keep_matrix`` (default True = reference behaviour) lets sweeps drop the N x N matrices;
* an exactly singular system raises ``numpy.linalg.LinAlgError`` for NumPy callers (as the reference's solve does); for
  torch callers its density is NaN instead (no host synchronisation inside a sweep).
"""

from __future__ import annotations

import contextlib
import math
import threading
import warnings
from collections.abc import Callable
from typing import Any, Literal, NotRequired, Protocol, TypedDict

import attrs
import numpy as np
import torch

from . import _ops
from ._coords import tree_spec
from ._lib import get_plan

Array = Any
F64 = torch.float64
C128 = torch.complex128


# --------------------------------------------------------------------------------------------------
# array-namespace plumbing
# --------------------------------------------------------------------------------------------------
class _NS:
    """Remembers the caller's array namespace so that results can be handed back in it."""

    def __init__(self, *arrays):
        self.kind = "numpy"
        self.device = None
        self.pinned = False
        for a in arrays:
            if isinstance(a, torch.Tensor):
                self.kind = "torch"
                self.device = a.device
                self.pinned = a.device.type == "cpu" and a.is_pinned()
                break

    def out(self, t: torch.Tensor):
        if self.kind == "numpy":
            t = t.detach()
            nbytes = t.numel() * t.element_size()
            if t.is_cuda and (1 << 20) <= nbytes <= (4 << 30):
                # large results (the N x N matrices the reference returns by default): one asynchronous copy into pinned
                # host memory from torch's caching host allocator instead of a pageable .cpu() (PCIe-rate, ~20x faster)
                host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                host.copy_(t, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return host.numpy()
            return t.cpu().numpy()
        if self.pinned and t.is_cuda:
            # pinned host tensors in -> pinned host tensors out (one asynchronous PCIe-rate copy)
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host.copy_(t.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return host
        return t.to(self.device)

    def asuser(self, t: torch.Tensor):
        return self.out(t)


def _dev() -> torch.device:
    return _ops._dev()


def _t(a, dtype=None) -> torch.Tensor:
    """To a torch tensor on the compute device (no copy when already there)."""
    if isinstance(a, torch.Tensor):
        t = a.to(_dev())
    else:
        t = torch.as_tensor(np.asarray(a), device=_dev())
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def _split_k(k: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor | None]:
    """(Re k, Im k) as float64 tensors; the imaginary part is None when it vanishes identically (real fast path)."""
    if k.is_complex():
        if bool(torch.any(k.imag != 0)):
            return k.real.to(F64), k.imag.to(F64)
        k = k.real
    return k.to(F64), None


def _real_k(k: torch.Tensor) -> torch.Tensor:
    kr, ki = _split_k(k)
    if ki is not None:
        raise NotImplementedError("complex wavenumbers (Im k != 0) are not implemented for this call")
    return kr


def harm_n_ndim_le(n_end: int, *, c_ndim: int) -> int:
    """Number of harmonics of degree < n_end on S^{c_ndim-1} (ush.harm_n_ndim_le, _biem.py:44)."""
    if c_ndim == 2:
        return max(2 * n_end - 1, 0)
    return sum(
        math.comb(n + c_ndim - 1, c_ndim - 1) - (math.comb(n + c_ndim - 3, c_ndim - 1) if n >= 2 else 0)
        for n in range(n_end)
    )


def max_memory(*, c_ndim: int, n_end: int, n_balls: int) -> int:
    """Maximum memory usage in bytes -- same model (and the same missing x16 for d <= 3) as _biem.py:23-49."""
    _COMPLEX128_SIZE = 16
    if c_ndim <= 3:
        return n_balls**2 * harm_n_ndim_le(n_end, c_ndim=c_ndim) ** 2

    def inner(c_ndim: int, n_end: int) -> int:
        return (2 * n_end - 1) * n_end ** (c_ndim - 1)

    return n_balls**2 * inner(c_ndim, n_end) ** 2 * inner(c_ndim, 2 * n_end) * _COMPLEX128_SIZE


def max_n_end(*, c_ndim: int, memory_limit: int, n_balls: int) -> int:
    """Maximum n_end that fits in the given memory limit (_biem.py:52-74)."""
    for i in range(1000):
        if max_memory(c_ndim=c_ndim, n_end=i, n_balls=n_balls) > memory_limit:
            break
    return i - 1


class BIEMKwargs(TypedDict):
    """The kwargs for the BIEM (_biem.py:77-101)."""

    centers: Array
    radii: Array
    k: Array
    n_end: int
    eta: NotRequired[Array]
    kind: NotRequired[Literal["inner", "outer"]]
    force_matrix: NotRequired[bool]


class UinCallable(Protocol):
    """Callable that computes the incident field at the given cartesian coordinates (_biem.py:104-128)."""

    def __call__(self, x: Array, /, *, expand_x: bool = True) -> Array: ...


class BIEMResultCalculatorProtocol(Protocol):
    """Structural type of the result (_biem.py:131-193)."""

    c: Any
    uin: UinCallable | None
    centers: Array
    radii: Array
    k: Array
    n_end: int
    eta: Array
    kind: Literal["inner", "outer"]
    density: Array | None
    matrix: Array | None

    def uscat(self, x: Array, /, far_field: bool = False, per_ball: bool = False, expand_x: bool = True) -> Array: ...


@attrs.frozen(kw_only=True)
class BIEMResultCalculator:
    """Result record of :func:`biem` (mirror of _biem.py:196-237).

    ``centers`` is stored transposed, ``[c_ndim, ..., B]``, exactly as the reference does (_biem.py:588,810).
    """

    c: Any
    uin: UinCallable | None = None
    centers: Array
    radii: Array
    k: Array
    n_end: int
    eta: Array
    kind: Literal["inner", "outer"]
    density: Array | None = None
    matrix: Array | None = None
    # device-resident copies of the fields above, attached by biem() for the evaluation kernels.  Not an __init__
    # argument: attrs.evolve(res, density=...) therefore builds a record WITHOUT it and biem_u falls back to the
    # public fields, as the reference's biem_u always does.
    _dev_state: dict = attrs.field(factory=dict, eq=False, repr=False, init=False)

    def uscat(self, x: Array, /, far_field: bool = False, per_ball: bool = False, expand_x: bool = True) -> Array:
        return biem_u(self, x, far_field=far_field, per_ball=per_ball, expand_x=expand_x)


# --------------------------------------------------------------------------------------------------
# input checks (mirror of _check_biem_inputs, _biem.py:240-326)
# --------------------------------------------------------------------------------------------------
def _check_biem_inputs(d: int, centers, radii, k, eta, alpha, beta):
    cen = _t(centers)
    rad = _t(radii)
    kk = _t(k)
    if eta is None:
        et = torch.ones((1,) * kk.dim(), dtype=F64, device=_dev())
    else:
        et = _t(eta)
    al = _t(alpha, C128)
    if al.dim() == 0:
        al = al[(None,) * (kk.dim() + 1)]
    be = _t(beta, C128)
    if be.dim() == 0:
        be = be[(None,) * (kk.dim() + 1)]
    if et.is_complex():
        raise ValueError("The decoupling parameter must be real.")
    et = et.to(F64)
    if bool(torch.any(et == 0)):
        warnings.warn(
            "The solution may be incorrect"
            "if k is an eigenvalue for laplacian"
            "on the interior region with"
            "Neumann boundary condition.",
            UserWarning,
            stacklevel=3,
        )
    kr = kk.real if kk.is_complex() else kk
    if bool(torch.any(et * kr < 0)) or (kk.is_complex() and bool(torch.any(kk.imag < 0))):
        warnings.warn(
            "The solution may be incorrectif not (Im k >= 0 and eta Re k >= 0).", UserWarning, stacklevel=3
        )
    if len({kk.dim(), et.dim(), cen.dim() - 2, rad.dim() - 1}) != 1:
        raise ValueError(
            f"{kk.dim()=}, {et.dim()=}, {cen.dim() - 2=}, {rad.dim() - 1=}are not the same."
        )
    try:
        torch.broadcast_shapes(kk.shape, et.shape, cen.shape[:-2], rad.shape[:-1], al.shape, be.shape)
    except Exception as e:
        raise ValueError(
            "Shapes of k, eta and centers.shape[:-2], radii.shape[:-1] are not broadcastable\n"
            f"{tuple(kk.shape)=}\n{tuple(et.shape)=}\n{tuple(cen.shape)=}\n{tuple(rad.shape)=}\n"
            f"{tuple(al.shape)=}\n{tuple(be.shape)=}"
        ) from e
    try:
        torch.broadcast_shapes(cen.shape[:-1], rad.shape, al.shape, be.shape)
    except Exception as e:
        raise ValueError(
            "centers.shape[:-1] and radii.shape are not broadcastable\n"
            f"{tuple(cen.shape)=}\n{tuple(rad.shape)=}\n{tuple(al.shape)=}\n{tuple(be.shape)=}"
        ) from e
    if cen.shape[-1] != d:
        raise ValueError(f"The last dimension of centers must be c.c_ndim={d}, but got {cen.shape[-1]}")
    return cen.to(F64), rad.to(F64), kk, et, al, be


# --------------------------------------------------------------------------------------------------
# incident fields (mirror of _biem.py:329-450)
# --------------------------------------------------------------------------------------------------
def _xp_of(*arrays):
    for a in arrays:
        if isinstance(a, torch.Tensor):
            return torch
    return np


def plane_wave(*, k: Array, direction: Array) -> tuple[Callable[[Array], Array], Callable[[Array], Array]]:
    r"""Plane wave ``u(x) = exp(i k d.x)``, ``d = direction / |direction|`` (_biem.py:329-388).

    The returned closures are ordinary array functions (NumPy or torch, following the inputs).  They also carry
    a ``_bhs_plane_wave`` tag so that :func:`biem` can evaluate the boundary data inside the fused
    right-hand-side kernel instead of calling back into Python.
    """
    xp = _xp_of(k, direction)
    if xp is np:
        k = np.asarray(k)
        direction = np.asarray(direction, dtype=np.float64)
    try:
        np.broadcast_shapes(tuple(k.shape), tuple(direction.shape[1:]))
    except Exception as e:
        raise ValueError(
            "Shapes of k and direction[1:] are not broadcastable\n"
            f"{tuple(k.shape)=}\n{tuple(direction.shape)=}"
        ) from e
    if direction.ndim != k.ndim + 1:
        raise ValueError(f"{direction.ndim=} is not {k.ndim + 1=}")
    if xp is np:
        direction = direction / np.linalg.norm(direction, axis=0, keepdims=True)
    else:
        direction = direction / torch.linalg.vector_norm(direction, dim=0, keepdim=True)

    def _ip(x):
        dd = direction[(slice(None),) + (None,) * (x.ndim - direction.ndim)]
        return (dd * x).sum(0), dd

    def inner(x: Array, /) -> Array:
        ip, _ = _ip(x)
        return xp.exp(1j * k * ip)

    def inner_grad(x: Array, /) -> Array:
        ip, dd = _ip(x)
        return 1j * k * dd * xp.exp(1j * k * ip)[None, ...]

    tag = {"k": k, "direction": direction}
    inner._bhs_plane_wave = tag  # type: ignore[attr-defined]
    inner_grad._bhs_plane_wave = tag  # type: ignore[attr-defined]
    return inner, inner_grad


def point_source(*, k: Array, source: Array, n: int) -> tuple[Callable[[Array], Array], Callable[[Array], Array]]:
    r"""Point source ``u(x) = h_n^{(d)}(k |x - source|)`` (_biem.py:391-450); Hankel values from bhs_bessel."""
    xp = _xp_of(k, source)
    if xp is np:
        k = np.asarray(k)
        source = np.asarray(source, dtype=np.float64)
    try:
        np.broadcast_shapes(tuple(k.shape), tuple(source.shape[1:]))
    except Exception as e:
        raise ValueError(
            f"Shapes of k and source[1:] are not broadcastable\n{tuple(k.shape)=}\n{tuple(source.shape)=}"
        ) from e
    if source.ndim != k.ndim + 1:
        raise ValueError(f"{source.ndim=} is not {k.ndim + 1=}")
    ns = _NS(k, source)

    def _hankel(z, d: int, derivative: bool):
        if z.is_complex() if isinstance(z, torch.Tensor) else np.iscomplexobj(z):
            return ns.out(_ops.bessel_z(d, 2, n, _t(z, C128), derivative)[..., n])
        zt = _t(z, F64)
        return ns.out(_ops.bessel(d, 2, n, zt, derivative)[..., n])

    def inner(x: Array, /) -> Array:
        xx = x - source[(slice(None),) + (None,) * (x.ndim - source.ndim)]
        r = (xx**2).sum(0) ** 0.5
        return _hankel(k * r, int(x.shape[0]), False)

    def inner_grad(x: Array, /) -> Array:
        xx = x - source[(slice(None),) + (None,) * (x.ndim - source.ndim)]
        r = (xx**2).sum(0) ** 0.5
        coeff = k * _hankel(k * r, int(x.shape[0]), True) / r
        return coeff[None, ...] * xx

    return inner, inner_grad


# --------------------------------------------------------------------------------------------------
# sweep engine: streams + CUDA graphs around (assemble -> LU solve) for many independent systems
# --------------------------------------------------------------------------------------------------
class _Slot:
    """Buffers of one group of `S` systems that are assembled and factorised in lock step."""

    def __init__(self, d: int, n_end: int, B: int, N: int, S: int, priority: int = 0):
        dev = _dev()
        plan = get_plan(d, n_end)
        self.stream = torch.cuda.Stream(device=dev, priority=priority)
        self.A = torch.empty((S, N, N), dtype=C128, device=dev)
        self.k = torch.ones((S,), dtype=F64, device=dev)
        self.k_im = torch.zeros((S,), dtype=F64, device=dev)
        self.eta = torch.ones((S,), dtype=F64, device=dev)
        self.rhs = torch.zeros((S, N), dtype=C128, device=dev)
        self.bufs = _ops.SolveBuffers(N, 1, S)
        self.work = _ops._work(_ops.load().bhs_assemble_workspace(plan.handle, B, S))
        self.graph: torch.cuda.CUDAGraph | None = None


class SweepEngine:
    """Assemble + solve many independent systems that share one geometry (different k / eta / rhs).

    Systems are processed in groups of ``batch``: one launch of every assembly / LU kernel handles the whole group
    (bhs_assemble with nsys = batch, bhs_zgesv_batched), which divides the number of launches -- and of CUDA-graph nodes
    -- per system by ``batch``.  ``nslots`` groups are in flight at a time, each on its own stream with its own buffers and
    a CUDA graph of the (assembly -> LU) sequence, so that the latency-bound panel steps of one group overlap the
    tensor-core updates of the others.  A last, partial group is padded with the slot's previous wavenumbers (their results
    are discarded).  Nothing here synchronises with the host.
    """

    def __init__(self, d: int, n_end: int, B: int, nslots: int = 3, use_graphs: bool = True, complex_k: bool = False,
                 batch: int = 1):
        self.d, self.n_end, self.B = d, n_end, B
        self.complex_k = complex_k
        self.batch = batch
        self.plan = get_plan(d, n_end)
        self.N = B * self.plan.H
        dev = _dev()
        self.cen = torch.zeros((B, d), dtype=F64, device=dev)
        self.rad = torch.ones((B,), dtype=F64, device=dev)
        self.al = torch.ones((B,), dtype=C128, device=dev)
        self.be = torch.zeros((B,), dtype=C128, device=dev)
        import os

        nprio = max(1, int(os.environ.get("BHS_SWEEP_PRIO", "1")))
        self.slots = [_Slot(d, n_end, B, self.N, batch, priority=-(i % nprio)) for i in range(nslots)]
        self.slot_bytes = nslots * batch * 16 * self.N * self.N
        # a CUDA graph buys nothing for one huge system (C5: N = 36 864, seconds per solve) and its warm-up pass would double it
        self.use_graphs = use_graphs and self.N <= 12000
        self.graphs_ready = False
        self.lock = threading.Lock()  # one sweep at a time per engine (slots, streams and graphs are shared state)

    def _assemble(self, s: _Slot) -> None:
        _ops.assemble(self.d, self.n_end, self.cen, self.rad, s.k, s.eta, self.al, self.be, out=s.A, work=s.work,
                      k_im=s.k_im if self.complex_k else None)

    def _body(self, s: _Slot) -> None:
        self._assemble(s)
        _ops.zgesv_batched_(s.A, s.rhs, s.bufs)

    def set_geometry(self, cen, rad, al, be) -> None:
        self.cen.copy_(cen)
        self.rad.copy_(rad)
        self.al.copy_(al)
        self.be.copy_(be)

    def _ensure_graphs(self) -> None:
        """One eager pass on slot 0 (loads every kernel and sets the function attributes), then the (assembly -> LU)
        sequence of EVERY slot is captured up front -- before any copy or solve of the sweep is enqueued, so that the
        device-wide synchronisations of graph capture never stall groups in flight.  `thread_local` capture mode: CUDA
        calls of other host threads (a pinning DataLoader thread, a second engine) do not invalidate the capture."""
        if self.graphs_ready or not self.use_graphs:
            return
        s0 = self.slots[0]
        with torch.cuda.stream(s0.stream):
            self._body(s0)
        s0.stream.synchronize()
        for s in self.slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s.stream, capture_error_mode="thread_local"):
                self._body(s)
            s.graph = g
        self.graphs_ready = True

    def run(self, ks, etas, f_hat, out_density, out_matrix=None, kis=None, out_info=None) -> None:
        """ks, etas (and kis = Im k for a complex_k engine): [K]; f_hat: [K, N]; out_density: [K, N]; out_matrix: [K, N, N]
        or None; out_info: int32 [K] or None (zero-pivot flag of every system, bhs_zgesv's `info`)."""
        with self.lock:
            self._run(ks, etas, f_hat, out_density, out_matrix, kis, out_info)

    def _run(self, ks, etas, f_hat, out_density, out_matrix, kis, out_info) -> None:
        K = ks.shape[0]
        S = self.batch
        if out_matrix is None:
            self._ensure_graphs()
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        for s in self.slots:
            s.stream.wait_event(ready)
        for g_i, i0 in enumerate(range(0, K, S)):
            n = min(S, K - i0)
            s = self.slots[g_i % len(self.slots)]
            with torch.cuda.stream(s.stream):
                s.k[:n].copy_(ks[i0 : i0 + n], non_blocking=True)
                if self.complex_k:
                    s.k_im[:n].copy_(kis[i0 : i0 + n], non_blocking=True)
                s.eta[:n].copy_(etas[i0 : i0 + n], non_blocking=True)
                s.rhs[:n].copy_(f_hat[i0 : i0 + n], non_blocking=True)
                if out_matrix is not None:
                    # the caller keeps the matrices: assemble, copy out, then factor the slot copies
                    self._assemble(s)
                    out_matrix[i0 : i0 + n].copy_(s.A[:n], non_blocking=True)
                    _ops.zgesv_batched_(s.A, s.rhs, s.bufs)
                elif s.graph is not None:
                    s.graph.replay()
                else:
                    self._body(s)
                out_density[i0 : i0 + n].copy_(s.rhs[:n], non_blocking=True)
                if out_info is not None:  # padded systems of a partial group are not reported
                    out_info[i0 : i0 + n].copy_(s.bufs.info[:n], non_blocking=True)
        for s in self.slots:
            done = torch.cuda.Event()
            done.record(s.stream)
            cur.wait_event(done)


_engines: dict = {}
_MAX_ENGINES = 4


def _sweep_shape(N: int, K: int) -> tuple[int, int]:
    """(batch, nslots) of a sweep of K systems with N unknowns each: `batch` systems per launch, `nslots` such groups in
    flight (each group owns a stream, batch N x N buffers and a CUDA graph).

    The LU panel steps are latency-bound, so throughput comes from keeping many systems in flight.  Measured on B200 at
    N = 4096 (systems/s): 1 x 3 / 1 x 12 / 1 x 32 (batch x groups) -> 54 / 106 / 142; 8 x 4 -> 134, 4 x 8 -> 138,
    2 x 16 -> 142, 8 x 8 -> 144, 4 x 16 -> 146, 2 x 32 -> 147: independent streams matter more than fewer launches.
    With the round-2 kernels (3M trailing update) the curve is flat at the top: 2 x 32 -> 165, 2 x 64 -> 165, 4 x 24 -> 164,
    4 x 32 -> 167 (tools/sweep_shapes.sh).  Defaults: 32 groups of 4 systems for long sweeps (128 systems in flight, 34 GB at
    N = 4096), groups of 2 below 64 systems (more independent streams for the same number in flight; tools/sweep_short.sh: 32
    systems 2 x 16 -> 170, 4 x 8 -> 168, 1 x 32 -> 151; 64 systems 2 x 32 -> 174, 4 x 16 -> 179); capped so that the
    buffers use at most a quarter of the free device memory.  Override with BHS_SWEEP_BATCH / BHS_SWEEP_SLOTS."""
    import os

    if K <= 1 or N > 12000:
        return 1, 1
    # cudaMemGetInfo costs ~20 ms of host time with the device idle: ask once per (device, N, K, overrides)
    key = (torch.cuda.current_device(), N, K, os.environ.get("BHS_SWEEP_BATCH"), os.environ.get("BHS_SWEEP_SLOTS"))
    hit = _shape_cache.get(key)
    if hit is not None:
        return hit
    _shape_cache[key] = shape = _sweep_shape_uncached(N, K)
    return shape


_shape_cache: dict = {}


def _sweep_shape_uncached(N: int, K: int) -> tuple[int, int]:
    import os

    batch = max(1, int(os.environ.get("BHS_SWEEP_BATCH", "4" if K >= 64 else "2")))
    batch = min(batch, K)
    free, _ = torch.cuda.mem_get_info()
    fit = max(1, int(free // 4 // (16 * N * N)))  # systems that fit
    batch = max(1, min(batch, fit))
    env = os.environ.get("BHS_SWEEP_SLOTS")
    nslots = max(1, int(env)) if env else 32
    nslots = max(1, min(nslots, -(-K // batch), fit // batch))
    return batch, nslots


def _get_engine(d: int, n_end: int, B: int, nslots: int, complex_k: bool = False, batch: int = 1) -> SweepEngine:
    key = (torch.cuda.current_device(), d, n_end, B, nslots, complex_k, batch)
    e = _engines.get(key)
    if e is None:
        # the slot matrices of cached engines stay allocated: keep the cache small (most recent first, at most
        # _MAX_ENGINES entries and never more than a quarter of the device memory in total)
        need = nslots * batch * 16 * (B * get_plan(d, n_end).H) ** 2
        _, total = torch.cuda.mem_get_info()
        while _engines and (len(_engines) >= _MAX_ENGINES or need + sum(x.slot_bytes for x in _engines.values()) > total // 4):
            _engines.pop(next(iter(_engines)))
        e = SweepEngine(d, n_end, B, nslots, complex_k=complex_k, batch=batch)
        _engines[key] = e
    else:
        _engines[key] = _engines.pop(key)  # move to the end: most recently used
    return e


_side_streams: dict = {}


def _uscat_streams(n: int = 8):
    dev = torch.cuda.current_device()
    if dev not in _side_streams:
        _side_streams[dev] = [torch.cuda.Stream(device=dev) for _ in range(n)]
    return _side_streams[dev]


def clear_engines() -> None:
    """Drop cached sweep engines (frees their N x N slot buffers)."""
    _engines.clear()
    _shape_cache.clear()


# --------------------------------------------------------------------------------------------------
# biem (mirror of _biem.py:453-819)
# --------------------------------------------------------------------------------------------------
def _boundary_data(spec, plan, ns, cen, rad, al, be, uin, uin_grad, batch_shape):
    """g[K, Q, B] from arbitrary callables (the reference's `f` closure, _biem.py:611-624).

    The callables see x of shape (c_ndim, Q, B, *batch) in the caller's namespace -- the reference's
    (c_ndim, ...(f), B, ...(first)) with the quadrature grid flattened -- and return (Q, B, *batch).
    ``cen`` and the plan's quadrature directions live in the chain frame of the tree; the callables get the points (and
    the normals) in the caller's cartesian frame.
    """
    d = spec.d
    nb = len(batch_shape)
    K = int(np.prod(batch_shape)) if nb else 1
    dirs_np, _ = plan.quadrature()
    dirs = torch.as_tensor(dirs_np, device=_dev())  # [d, Q]
    Q = dirs.shape[1]
    B = int(torch.broadcast_shapes(cen.shape[-2:-1], rad.shape[-1:])[0])
    cen_e = cen.expand(tuple(batch_shape) + (B, d))
    rad_e = rad.expand(tuple(batch_shape) + (B,))
    x = (
        rad_e.movedim(-1, 0)[None, None] * dirs[(...,) + (None,) * (1 + nb)]
        + cen_e.movedim(-1, 0).movedim(-1, 1)[:, None]
    )  # [d, Q, B, *batch]
    if not spec.identity:
        inv = list(spec.inverse)
        x, dirs = x[inv], dirs[inv]
    xu = ns.out(x.contiguous())
    yhat = ns.out(dirs)[(...,) + (None,) * (1 + nb)]  # [d, Q, 1, *1]

    def coef(a):  # [..., B] -> [1, B, *batch-ish]
        while a.dim() < nb + 1:
            a = a[None]
        return ns.out(a.movedim(-1, 0)[None])

    g = 0
    if uin is not None:
        g = g - coef(al) * uin(xu)
    if uin_grad is not None:
        g = g - coef(be) * (uin_grad(xu) * yhat).sum(0)
    gt = _t(g, C128)
    gt = torch.broadcast_to(gt, (Q, B) + tuple(batch_shape))
    return gt.reshape(Q, B, K).permute(2, 0, 1).contiguous()


def biem(
    c: Any,
    /,
    *,
    centers: Array,
    radii: Array,
    k: Array,
    n_end: int,
    alpha: Array | complex = 1.0,
    beta: Array | complex = 0.0,
    uin: Callable[[Array], Array] | None = None,
    uin_grad: Callable[[Array], Array] | None = None,
    eta: Array | None = None,
    kind: Literal["inner", "outer"] = "outer",
    force_matrix: bool = False,
    translational_coefficients_method: Literal["gumerov", "plane_wave", "triplet"] | None = None,
    keep_matrix: bool = True,
) -> BIEMResultCalculator:
    r"""Boundary integral equation method for the Helmholtz equation around non-overlapping n-spheres.

    Drop-in for the reference ``biem`` (_biem.py:453-819): RHS expansion (bhs_rhs_expand), assembly of the
    diagonal blocks and the (S|R) translation blocks (bhs_assemble), dense complex128 solve (bhs_zgesv) -- or the
    single-sphere shortcut (bhs_diag_coef) -- and packaging into a :class:`BIEMResultCalculator`.

    ``translational_coefficients_method`` is accepted for signature compatibility; the B200 path always uses the
    per-entry-exact sparse coupling sum (no `triplet` quadrature noise).
    """
    spec = tree_spec(c)
    d = spec.d
    ns = _NS(centers, radii, k, eta)
    cen_user, rad, kk, et, al, be = _check_biem_inputs(d, centers, radii, k, eta, alpha, beta)
    # the device code works in the chain frame of the tree (b' / relabelled trees: a permutation of the cartesian axes)
    cen = cen_user if spec.identity else cen_user[..., list(spec.axes)].contiguous()
    del translational_coefficients_method
    nb = kk.dim()
    batch_shape = tuple(torch.broadcast_shapes(kk.shape, et.shape, cen.shape[:-2], rad.shape[:-1], al.shape[:-1], be.shape[:-1]))
    K = int(np.prod(batch_shape)) if nb else 1
    B = int(torch.broadcast_shapes(cen.shape[-2:-1], rad.shape[-1:], al.shape[-1:], be.shape[-1:])[0])
    plan = get_plan(d, n_end)
    H = plan.H
    N = B * H

    shared_geom = all(int(np.prod(t.shape[:-1])) == 1 for t in (rad, al, be)) and int(np.prod(cen.shape[:-2])) == 1
    k_user = kk
    kk, kk_im = _split_k(kk)
    ks = kk.expand(batch_shape).reshape(K).contiguous()
    kis = None if kk_im is None else kk_im.expand(batch_shape).reshape(K).contiguous()
    ets = et.expand(batch_shape).reshape(K).contiguous()

    def geom(i):
        def pick(t, tail):
            tb = t.expand(batch_shape + tuple(t.shape[-tail:])) if nb else t
            return tb.reshape((K,) + tuple(t.shape[-tail:]))[i]

        c_i = pick(cen, 2).expand(B, d).contiguous()
        r_i = pick(rad, 1).expand(B).contiguous()
        a_i = pick(al, 1).expand(B).contiguous()
        b_i = pick(be, 1).expand(B).contiguous()
        return c_i, r_i, a_i, b_i

    # ---- right-hand side --------------------------------------------------------------------------
    f_hat = None
    if uin is not None or uin_grad is not None:
        if not bool(torch.all(al == 0)) and uin is None:
            raise ValueError("alpha is not zero, but uin is None. uin must be provided to compute the boundary condition.")
        if not bool(torch.all(be == 0)) and uin_grad is None:
            raise ValueError("beta is not zero, but uin_grad is None. uin_grad must be provided to compute the boundary condition.")
        tag = getattr(uin if uin is not None else uin_grad, "_bhs_plane_wave", None)
        fused = (
            tag is not None
            and shared_geom
            and (uin is None or getattr(uin, "_bhs_plane_wave", None) is tag)
            and (uin_grad is None or getattr(uin_grad, "_bhs_plane_wave", None) is tag)
            and np.ndim(tag["direction"]) >= 1
            and int(np.prod(tuple(tag["direction"].shape[1:]))) == 1
        )
        if fused:
            c0, r0, a0, b0 = geom(0)
            kin, kin_im = _split_k(_t(tag["k"]))
            kin = kin.expand(batch_shape).reshape(K).contiguous() if nb else kin.reshape(1)
            if kin_im is not None:
                kin_im = kin_im.expand(batch_shape).reshape(K).contiguous() if nb else kin_im.reshape(1)
            dirv = _t(tag["direction"], F64).reshape(d)[list(spec.axes)].contiguous()
            f_hat = _ops.rhs_expand(d, n_end, centers=c0, radii=r0, k_in=kin, direction=dirv,
                                    alpha=a0 if uin is not None else torch.zeros_like(a0),
                                    beta=b0 if uin_grad is not None else None, k_in_im=kin_im, tree=spec.tree)
        else:
            g = _boundary_data(spec, get_plan(d, n_end, spec.tree), ns, cen, rad, al, be, uin, uin_grad, batch_shape)
            f_hat = _ops.rhs_expand(d, n_end, g=g, tree=spec.tree)  # [K, B, H]

    use_matrix = (uin is None and uin_grad is None) or B > 1 or force_matrix

    density_t = None
    matrix_t = None
    if not use_matrix:
        # single-sphere shortcut (_biem.py:648-691)
        dens = []
        for i in range(K) if not shared_geom else [None]:
            if i is None:
                c0, r0, a0, b0 = geom(0)
                diag = _ops.diag_coef(d, n_end, r0, ks, ets, a0, b0, k_im=kis)  # [K, B, H]
                dens = f_hat / diag
            else:
                c_i, r_i, a_i, b_i = geom(i)
                dens.append(f_hat[i] / _ops.diag_coef(d, n_end, r_i, ks[i : i + 1], ets[i : i + 1], a_i, b_i,
                                                      k_im=None if kis is None else kis[i : i + 1])[0])
        density_t = dens if isinstance(dens, torch.Tensor) else torch.stack(dens)
    else:
        if f_hat is None:
            # matrix only (_biem.py:596-597,794-795)
            if shared_geom:
                c0, r0, a0, b0 = geom(0)
                matrix_t = _ops.assemble(d, n_end, c0, r0, ks, ets, a0, b0, k_im=kis)
            else:
                matrix_t = torch.stack([
                    _ops.assemble(d, n_end, *geom(i)[:2], ks[i : i + 1], ets[i : i + 1], *geom(i)[2:],
                                  k_im=None if kis is None else kis[i : i + 1])[0] for i in range(K)
                ])
        else:
            density_t = torch.empty((K, N), dtype=C128, device=_dev())
            matrix_t = torch.empty((K, N, N), dtype=C128, device=_dev()) if keep_matrix else None
            info_t = torch.zeros((K,), dtype=torch.int32, device=_dev())
            rhs = f_hat.reshape(K, N)
            if shared_geom:
                batch, nslots = _sweep_shape(N, K)
                eng = _get_engine(d, n_end, B, nslots, complex_k=kis is not None, batch=batch)
                eng.set_geometry(*geom(0))
                eng.run(ks, ets, rhs, density_t, matrix_t, kis=kis, out_info=info_t)
            else:
                eng = _get_engine(d, n_end, B, 1, complex_k=kis is not None)
                for i in range(K):
                    eng.set_geometry(*geom(i))
                    eng.run(ks[i : i + 1], ets[i : i + 1], rhs[i : i + 1], density_t[i : i + 1],
                            None if matrix_t is None else matrix_t[i : i + 1],
                            kis=None if kis is None else kis[i : i + 1], out_info=info_t[i : i + 1])
            # Exactly singular systems (zero pivot, bhs_zgesv `info` != 0): the reference's solve raises LinAlgError.  NumPy
            # callers get the same (their results are copied to the host anyway); for torch callers the check stays on the
            # device -- the densities of the singular systems become NaN, so they cannot be mistaken for solutions, and a
            # sweep is never synchronised with the host.
            if ns.kind == "numpy":
                bad = torch.nonzero(info_t).reshape(-1)
                if bad.numel():
                    i_bad = int(bad[0])
                    raise np.linalg.LinAlgError(
                        f"Singular matrix: zero pivot at step {int(info_t[i_bad]) - 1} of system {i_bad} (k = {complex(ks[i_bad]) if kis is None else complex(ks[i_bad], kis[i_bad])})")
            else:
                density_t = torch.where((info_t != 0)[:, None], torch.full_like(density_t, float("nan")), density_t)

    # ---- packaging (user namespace) ---------------------------------------------------------------
    density = None if density_t is None else ns.out(density_t.reshape(batch_shape + (B, H)))
    matrix = None if matrix_t is None else ns.out(matrix_t.reshape(batch_shape + (B, H, B, H)))
    ndim_first = nb

    if uin is None:
        uin_wrapped = None
    else:

        def uin_wrapped(x: Array, /, *, expand_x: bool = True) -> Array:
            if expand_x:
                x = x[(...,) + (None,) * ndim_first]
            return uin(x)

    cen_store = ns.out(torch.movedim(cen_user, -1, 0))  # [v, ..., B]  (_biem.py:588)
    dev_state = {
        "bt": spec.chain, "batch_shape": batch_shape, "K": K, "B": B, "ks": ks, "etas": ets,
        "ks_host": ks.tolist() if kis is None else torch.complex(ks, kis).tolist(), "etas_host": ets.tolist(),
        "cen": cen, "rad": rad, "density": density_t, "geom_shared": shared_geom,
    }
    res = BIEMResultCalculator(
        c=c, centers=cen_store, radii=ns.out(rad), k=ns.out(k_user), n_end=n_end, eta=ns.out(et), kind=kind,
        uin=uin_wrapped, density=density, matrix=matrix,
    )
    object.__setattr__(res, "_dev_state", dev_state)
    return res


# --------------------------------------------------------------------------------------------------
# biem_u (mirror of _biem.py:822-977)
# --------------------------------------------------------------------------------------------------
_PIPE_MIN_POINTS = 1 << 19
_pipe_streams: dict = {}


def _uscat_host_pipelined(spec, d, n_end, B, cen, rad, st, dens, x, far_field, per_ball, inner, ns):
    """Field of ONE system at host points ``x`` [d, ...]: returns the host result, or None when the point set is too small
    (or not float64 host memory) for the tiled path to pay."""
    if isinstance(x, torch.Tensor):
        xh = x
    else:
        xa = np.asarray(x)
        if xa.dtype != np.float64:
            return None
        xh = torch.from_numpy(np.ascontiguousarray(xa))
    if xh.dtype != F64 or xh.dim() < 2 or xh.shape[0] != d:
        return None
    xshape = tuple(xh.shape[1:])
    xh = xh.reshape(d, -1)
    P = xh.shape[1]
    if P < _PIPE_MIN_POINTS or xh.stride(1) != 1:
        return None
    dev = _dev()
    key = torch.cuda.current_device()
    if key not in _pipe_streams:
        _pipe_streams[key] = tuple(torch.cuda.Stream(device=dev) for _ in range(3))
    s_in, s_k, s_out = _pipe_streams[key]
    k = (st.get("ks_host") or st["ks"].tolist())[0]
    eta = (st.get("etas_host") or st["etas"].tolist())[0]
    cen, rad, dens = cen.contiguous(), rad.contiguous(), dens.contiguous()
    tail = (B,) if per_ball else ()
    host = torch.empty((P,) + tail, dtype=C128, pin_memory=True)
    ntile = max(2, min(16, P // (1 << 19)))
    bounds = [P * i // ntile for i in range(ntile + 1)]
    cur = torch.cuda.current_stream()
    ready = torch.cuda.Event()
    ready.record(cur)
    for s_ in (s_in, s_k, s_out):
        s_.wait_event(ready)
    work = _ops._work(_ops.load().bhs_uscat_workspace(get_plan(d, n_end).handle, B))
    keep = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        with torch.cuda.stream(s_in):
            xd = torch.empty((d, hi - lo), dtype=F64, device=dev)
            for i in range(d):  # chain frame of the tree; every row is one contiguous (pinned: asynchronous) copy
                xd[i].copy_(xh[spec.axes[i], lo:hi], non_blocking=True)
            e_in = torch.cuda.Event()
            e_in.record(s_in)
        with torch.cuda.stream(s_k):
            s_k.wait_event(e_in)
            od = _ops.uscat(d, n_end, cen, rad, k, eta, dens, xd, far_field=far_field, per_ball=per_ball, inner=inner,
                            work=work)
            e_k = torch.cuda.Event()
            e_k.record(s_k)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e_k)
            host[lo:hi].copy_(od, non_blocking=True)
        keep.append((xd, od))
    s_out.synchronize()
    s_k.synchronize()
    del keep
    host = host.reshape(xshape + tail)
    if ns.kind == "numpy":
        return host.numpy()
    return host  # torch CPU callers get the pinned host tensor itself

def biem_u(res: Any, x: Array, /, far_field: bool = False, per_ball: bool = False, expand_x: bool = True) -> Array:
    """Scattered field at the cartesian points ``x`` of shape ``(c_ndim, ...(x))`` (``+ ...(first)`` when
    ``expand_x`` is False); returns ``(...(x), ...(first))`` (``+ (B,)`` with ``per_ball``)."""
    if res.density is None:
        raise ValueError("The BIEMResult does not have density.")
    if res.kind not in ("outer", "inner"):
        raise ValueError(f"Invalid kind: {res.kind}")
    spec = tree_spec(res.c)
    d = spec.d
    ns = _NS(x, res.centers, res.radii, res.k, res.eta)
    st = getattr(res, "_dev_state", None) or {}
    if not st:
        # a result record built by hand (or by another implementation): move its fields to the device
        kk, kk_im = _split_k(_t(res.k))
        cen = torch.movedim(_t(res.centers, F64), 0, -1)
        if not spec.identity:
            cen = cen[..., list(spec.axes)]  # chain frame
        rad = _t(res.radii, F64)
        et = _t(res.eta, F64)
        dens = _t(res.density, C128)
        batch_shape = tuple(kk.shape)
        K = int(np.prod(batch_shape)) if batch_shape else 1
        B = dens.shape[-2]
        st = {
            "batch_shape": batch_shape, "K": K, "B": B,
            "ks": kk.expand(batch_shape).reshape(K),
            "ks_host": (kk.expand(batch_shape).reshape(K).tolist() if kk_im is None else
                        torch.complex(kk, kk_im).expand(batch_shape).reshape(K).tolist()),
            "etas": et.expand(batch_shape).reshape(K) if et.numel() > 1 else et.reshape(1).expand(K),
            "cen": cen, "rad": rad, "density": dens.reshape(K, -1),
        }
    batch_shape, K, B = st["batch_shape"], st["K"], st["B"]
    nb = len(batch_shape)
    n_end = res.n_end
    H = get_plan(d, n_end).H
    dens = st["density"].reshape(K, B, H)
    cen, rad = st["cen"], st["rad"]
    cen_k = cen.expand(batch_shape + (B, d)).reshape(K, B, d) if nb else cen.reshape(1, B, d)
    rad_k = rad.expand(batch_shape + (B,)).reshape(K, B) if nb else rad.reshape(1, B)

    # One solved system evaluated on a large HOST point set (a heat map): tiles of points travel host -> device -> host on
    # three streams, so that the PCIe copies of neighbouring tiles hide behind the field kernel.
    if nb == 0 and not (isinstance(x, torch.Tensor) and x.is_cuda):
        piped = _uscat_host_pipelined(spec, d, n_end, B, cen_k[0], rad_k[0], st, dens[0], x, far_field, per_ball,
                                      res.kind == "inner", ns)
        if piped is not None:
            return piped
    xs = torch.stack([_t(x[spec.axes[i]], F64) for i in range(d)], dim=0)  # chain frame
    if expand_x or nb == 0:
        xshape = tuple(xs.shape[1:])
        xf = xs.reshape(d, -1).contiguous()
        per_sys = [xf] * K
    else:
        xshape = tuple(xs.shape[1 : xs.dim() - nb])
        xb = torch.broadcast_to(xs, (d,) + xshape + batch_shape).reshape(d, -1, K)
        per_sys = [xb[:, :, i].contiguous() for i in range(K)]
    ks_host = st.get("ks_host") or st["ks"].tolist()
    etas_host = st.get("etas_host") or st["etas"].tolist()
    outs = []
    P_pts = per_sys[0].shape[1] if K else 0
    # Small point sets (probe points of a sweep) are latency-bound, one CTA each: spread the systems over a few side
    # streams so that the launches overlap.  Large point sets fill the device on their own.
    side = _uscat_streams() if (K > 4 and P_pts <= 8192) else None
    if side:
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        for s_ in side:
            s_.wait_event(ready)
    for i in range(K):
        ctx = torch.cuda.stream(side[i % len(side)]) if side else contextlib.nullcontext()
        with ctx:
            o = _ops.uscat(d, n_end, cen_k[i].contiguous(), rad_k[i].contiguous(), ks_host[i], etas_host[i],
                           dens[i].contiguous(), per_sys[i], far_field=far_field, per_ball=per_ball,
                           inner=(res.kind == "inner"))
        if side:
            o.record_stream(cur)
        outs.append(o)
    if side:
        for s_ in side:
            done = torch.cuda.Event()
            done.record(s_)
            cur.wait_event(done)
    out = torch.stack(outs, dim=1) if nb else outs[0]  # [P, K(, B)] or [P(, B)]
    tail = (B,) if per_ball else ()
    out = out.reshape(xshape + batch_shape + tail)
    return ns.out(out)
