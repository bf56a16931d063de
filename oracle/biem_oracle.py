"""NumPy/SciPy float64 oracle of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Restates, for chain coordinate types ``"a"`` (2-D), ``"ba"`` (3-D), ``"bba"`` (4-D)
(and deeper ``b...ba`` chains), what

* ``biem``       -- /root/reference/src/biem_helmholtz_sphere/_biem.py:453-819
* ``biem_u``     -- _biem.py:822-977
* ``plane_wave`` -- _biem.py:329-388,  ``point_source`` -- _biem.py:391-450

compute through the un-vendored packages ``ultrasphere==2.0.4``,
``ultrasphere-harmonics==1.3.0`` and ``batch-tensorsolve==1.0.1`` (pinned in the
reference's ``uv.lock``; absent from this container).  The published algorithm of
those packages is restated from SURVEY.md Appendix A; parity is pinned on the
reference's golden CSVs (tests/golden/, tests/test_oracle_golden.py).

Nothing in the product package imports this module.
"""

from __future__ import annotations

import functools
import math
import re
from dataclasses import dataclass
from typing import Callable

import numpy as np
import scipy.sparse
from scipy import special as sp

# --------------------------------------------------------------------------------------
# coordinates  (stand-in for ultrasphere.SphericalCoordinates)
# --------------------------------------------------------------------------------------


_TOKEN = re.compile(r"bp|b'|a|b|c")


def _parse_tree(branching_types: str):
    """(node types, hopf?) of a branching-types string: chains of b / b' ('bp') nodes over one a node, or 'caa'."""
    pos, toks = 0, []
    while pos < len(branching_types):
        m = _TOKEN.match(branching_types, pos)
        if not m:
            raise ValueError(f"invalid branching types {branching_types!r}")
        toks.append("bp" if m.group(0) in ("bp", "b'") else m.group(0))
        pos = m.end()
    if toks == ["c", "a", "a"]:
        return tuple(toks), True
    if not toks or toks[-1] != "a" or set(toks[:-1]) - {"b", "bp"}:
        raise ValueError(f"oracle supports chains of b / bp nodes over an a node and 'caa': got {branching_types!r}")
    return tuple(toks), False


class OracleCoordinates:
    """Coordinate trees of the reference's sweeps (cli.py:41): chains ``'a'``, ``'ba'``, ``'bpa'``, ``'bba'``, ``'bpbpa'``, ...
    and the Hopf tree ``'caa'``.

    type b : leaf = r cos t, sub-tree = r sin t, t in [0, pi];   type b' ('bp'): leaf = r sin t, sub-tree = r cos t,
    t in [-pi/2, pi/2];   'caa': x = r (cos t0 cos t1, cos t0 sin t1, sin t0 cos t2, sin t0 sin t2), t0 in [0, pi/2]
    (decoded from the reference's *.svg drawings; pinned by the golden rows of jascome_output.csv).
    ``axes``: tree coordinate i is the caller's cartesian coordinate axes[i] (``relabel`` = the nx.relabel_nodes of
    cli.py:63-69).  Every tree of one dimension spans the same harmonic space; the tables below (index tables, harmonics,
    coupling coefficients) are those of the plain chain ``.chain`` in the frame ``axes`` -- what the tree changes is where
    the right-hand-side quadrature samples the sphere.
    """

    def __init__(self, branching_types: str, axes=None):
        self.nodes, self.hopf = _parse_tree(branching_types)
        self.branching_types_expression_str = branching_types
        self.s_ndim = len(self.nodes)
        self.c_ndim = self.s_ndim + 1
        self.chain = "b" * (self.c_ndim - 2) + "a"
        self.root = 0
        self.axes = tuple(range(self.c_ndim)) if axes is None else tuple(int(a) for a in axes)
        assert sorted(self.axes) == list(range(self.c_ndim))

    def relabel(self, mapping: dict) -> "OracleCoordinates":
        return OracleCoordinates(self.branching_types_expression_str, tuple(mapping.get(a, a) for a in self.axes))

    @property
    def inverse(self):
        return tuple(int(i) for i in np.argsort(self.axes))

    def to_cartesian(self, spherical, as_array: bool = True):
        d = self.c_ndim
        r = spherical.get("r", 1.0)
        if self.hopf:
            c0, s0 = np.cos(spherical[0]), np.sin(spherical[0])
            y = [r * c0 * np.cos(spherical[1]), r * c0 * np.sin(spherical[1]),
                 r * s0 * np.cos(spherical[2]), r * s0 * np.sin(spherical[2])]
        else:
            y = []
            prod = r
            for i in range(d - 2):
                if self.nodes[i] == "b":
                    y.append(prod * np.cos(spherical[i]))
                    prod = prod * np.sin(spherical[i])
                else:
                    y.append(prod * np.sin(spherical[i]))
                    prod = prod * np.cos(spherical[i])
            y.append(prod * np.cos(spherical[d - 2]))
            y.append(prod * np.sin(spherical[d - 2]))
        out = [None] * d
        for i in range(d):
            out[self.axes[i]] = y[i]
        return np.stack(np.broadcast_arrays(*out), axis=0) if as_array else dict(enumerate(out))

    def from_cartesian(self, x):
        d = self.c_ndim
        y = [np.asarray(x[self.axes[i]], dtype=np.float64) for i in range(d)]
        if self.hopf:
            ra, rb = np.hypot(y[0], y[1]), np.hypot(y[2], y[3])
            return {"r": np.hypot(ra, rb), 0: np.arctan2(rb, ra), 1: np.arctan2(y[1], y[0]), 2: np.arctan2(y[3], y[2])}
        out = chain_from_cartesian(y)
        for i in range(d - 2):
            if self.nodes[i] == "bp":
                out[i] = np.pi / 2 - out[i]
        return out


def chain_from_cartesian(y):
    """Angles of the plain chain ``b...ba``: y0 = r cos t0, y1 = r sin t0 cos t1, ..., y_{d-1} = r sin t0 ... sin t_{d-2}."""
    d = len(y)
    y = [np.asarray(v, dtype=np.float64) for v in y]
    out = {}
    tail = np.zeros_like(y[0])
    tails = [None] * d
    for i in range(d - 1, -1, -1):
        tail = tail + y[i] ** 2
        tails[i] = tail
    out["r"] = np.sqrt(tails[0])
    for i in range(d - 2):
        out[i] = np.arctan2(np.sqrt(tails[i + 1]), y[i])
    out[d - 2] = np.arctan2(y[d - 1], y[d - 2])
    return out


def chain_to_cartesian(angles, d):
    out = []
    prod = 1.0
    for i in range(d - 1):
        out.append(prod * np.cos(angles[i]))
        prod = prod * np.sin(angles[i])
    out.append(prod)
    return np.stack(np.broadcast_arrays(*out), axis=0)


def create_from_branching_types(branching_types: str) -> OracleCoordinates:
    return OracleCoordinates(branching_types)


def _coords(c) -> OracleCoordinates:
    return c if isinstance(c, OracleCoordinates) else OracleCoordinates(
        c if isinstance(c, str) else c.branching_types_expression_str, getattr(c, "axes", None))


def _btype(c) -> str:
    """Chain tree whose tables serve ``c`` ('a', 'ba', 'bba', ...)."""
    if isinstance(c, str) and re.fullmatch(r"b*a", c):
        return c
    return _coords(c).chain


# --------------------------------------------------------------------------------------
# radial functions  z_n^{(d)}(x) = sqrt(pi/2) Z_{n+d/2-1}(x) / x^{d/2-1}   
# --------------------------------------------------------------------------------------


def radial(d: int, n_max: int, z, kind: str = "j", derivative: bool = False):
    """Return array [n_max+1, *z.shape] of j/y/h1 type hyperspherical functions (or d/dz)."""
    z = np.asarray(z)
    n = np.arange(n_max + 2).reshape((-1,) + (1,) * z.ndim)
    nu = n + d / 2.0 - 1.0
    fn = {"j": sp.jv, "y": sp.yv, "h": sp.hankel1}[kind]
    val = math.sqrt(math.pi / 2.0) * fn(nu, z[None]) / z[None] ** (d / 2.0 - 1.0)
    if not derivative:
        return val[:-1]
    return (n[:-1] / z[None]) * val[:-1] - val[1:]


# --------------------------------------------------------------------------------------
# index tables and harmonics   
# --------------------------------------------------------------------------------------


def harm_count(d: int, n_end: int) -> int:
    if d == 2:
        return 2 * n_end - 1
    return sum(math.comb(n + d - 1, d - 1) - (math.comb(n + d - 3, d - 1) if n >= 2 else 0) for n in range(n_end))


@functools.lru_cache(maxsize=None)
def index_tables(btype: str, n_end: int):
    """Flattened harmonic index table, [H, s_ndim] int: columns (n_0, n_1, ..., m_signed).

    Order: C-order over the unflattened axes (b-axes 0..n_end-1, a-axis in FFT order
    0..L-1, -(L-1)..-1) restricted to n_0 >= n_1 >= ... >= |m|   ([RECALL] ordering of
    ush.flatten_harmonics; unpinned by any reference artefact).
    """
    s = len(btype)
    ms = list(range(n_end)) + list(range(-(n_end - 1), 0))
    rows = []

    def rec(prefix, upper, depth):
        if depth == s - 1:
            for m in ms:
                if abs(m) <= upper:
                    rows.append(prefix + [m])
            return
        for v in range(0, upper + 1):
            rec(prefix + [v], v, depth + 1)

    rec([], n_end - 1, 0)
    tab = np.asarray(rows, dtype=np.int64).reshape(-1, s)
    assert tab.shape[0] == harm_count(s + 1, n_end)
    return tab


def degree_table(btype: str, n_end: int) -> np.ndarray:
    tab = index_tables(btype, n_end)
    return np.abs(tab[:, 0]) if len(btype) == 1 else tab[:, 0]


def _node_function(s: int, n, l, theta):
    """Normalised f^{(s)}_{n,l}(theta) = A sin^l C^{(l+s/2)}_{n-l}(cos), int f^2 sin^s = 1."""
    n = np.asarray(n)
    l = np.asarray(l)
    lam = l + s / 2.0
    kk = n - l
    lognorm = (
        math.log(math.pi)
        + (1.0 - 2.0 * lam) * math.log(2.0)
        + sp.gammaln(kk + 2.0 * lam)
        - sp.gammaln(kk + 1.0)
        - np.log(kk + lam)
        - 2.0 * sp.gammaln(lam)
    )
    x = np.cos(theta)
    st = np.sin(theta)
    with np.errstate(divide="ignore", invalid="ignore"):
        sl = np.where(l == 0, 1.0, st ** l)
    return np.exp(-0.5 * lognorm) * sl * sp.eval_gegenbauer(kk, lam, x)


def harmonics(c, angles, n_end: int) -> np.ndarray:
    """Orthonormal harmonics Y_h at the given angles -> [..., H]  (Phase(0), ultrasphere order)."""
    btype = _btype(c)
    s = len(btype)
    tab = index_tables(btype, n_end)
    th = [np.asarray(angles[i], dtype=np.float64)[..., None] for i in range(s)]
    m = tab[:, s - 1]
    out = np.exp(1j * m * th[s - 1]) / math.sqrt(2.0 * math.pi)
    for i in range(s - 1):
        nn = tab[:, i]
        ll = np.abs(tab[:, i + 1])
        out = out * _node_function(s - 1 - i, nn, ll, th[i])
    return out


# --------------------------------------------------------------------------------------
# quadrature and expansion   (under-resolved product rule, n = n_end)
# --------------------------------------------------------------------------------------


@functools.lru_cache(maxsize=None)
def quadrature(btype: str, n: int):
    """Product rule: angles list (each [Q]) and weights [Q]; node 0 is the slowest axis."""
    s = len(btype)
    axes, wts = [], []
    for i in range(s - 1):
        desc = s - 1 - i
        a = (desc - 1) / 2.0
        x, w = sp.roots_jacobi(n, a, a)
        axes.append(np.arccos(x))
        wts.append(w)
    axes.append(2.0 * math.pi * np.arange(2 * n) / (2 * n))
    wts.append(np.full(2 * n, math.pi / n))
    grids = np.meshgrid(*axes, indexing="ij")
    wg = np.meshgrid(*wts, indexing="ij")
    w = np.ones_like(wg[0])
    for ww in wg:
        w = w * ww
    return [g.ravel() for g in grids], w.ravel()


@functools.lru_cache(maxsize=None)
def _hopf_quadrature(n: int):
    """Product rule of the 'caa' tree: n Gauss-Legendre nodes in cos 2 t0 (surface measure sin t0 cos t0 dt0 dt1 dt2 =
    d(cos 2 t0)/4 dt1 dt2), 2n equispaced nodes on each circle.  Returns unit vectors [4, Q] and weights [Q]."""
    t, w0 = sp.roots_legendre(n)
    th0 = 0.5 * np.arccos(t)
    az = 2.0 * math.pi * np.arange(2 * n) / (2 * n)
    T0, T1, T2 = (a.ravel() for a in np.meshgrid(th0, az, az, indexing="ij"))
    W = (0.25 * w0[:, None, None] * np.full((1, 2 * n, 1), math.pi / n) * np.full((1, 1, 2 * n), math.pi / n)).ravel()
    y = np.stack([np.cos(T0) * np.cos(T1), np.cos(T0) * np.sin(T1), np.sin(T0) * np.cos(T2), np.sin(T0) * np.sin(T2)])
    return y, W


def quadrature_nodes(c, n: int):
    """Right-hand-side quadrature of the tree of ``c`` in its chain frame: (unit vectors [d, Q], chain angles, weights)."""
    co = _coords(c)
    if co.hopf:
        y, w = _hopf_quadrature(n)
        sph = chain_from_cartesian(list(y))
        return y, [sph[i] for i in range(co.c_ndim - 1)], w
    angles, w = quadrature(co.chain, n)  # b' nodes: theta' = pi/2 - theta, the same points
    return chain_to_cartesian(angles, co.c_ndim), angles, w


def expand(c, g_values: np.ndarray, n_end: int) -> np.ndarray:
    """f_hat[..., h] = sum_q w_q g[q, ...] conj Y_h(y_q)   (ush.expand, _biem.py:627)."""
    _, angles, w = quadrature_nodes(c, n_end)
    Y = harmonics(_btype(c), angles, n_end)  # [Q, H]
    return np.einsum("q,q...,qh->...h", w, g_values, np.conj(Y))


# --------------------------------------------------------------------------------------
# translation coefficients (S|R)   (sparse, structurally non-zero triples only)
# --------------------------------------------------------------------------------------


def _node_triple_tables(s: int, n_end: int):
    """Quadrature nodes/weights and F[n, l, q] (n < 2*n_end-1) for a b-node with s descendants."""
    nmax = 2 * n_end - 1
    nq = 2 * n_end + 2
    a = (s - 1) / 2.0
    x, w = sp.roots_jacobi(nq, a, a)
    theta = np.arccos(x)
    F = np.zeros((nmax, nmax, nq))
    for l in range(nmax):
        ns = np.arange(l, nmax)
        F[l:, l, :] = _node_function(s, ns[:, None], l, theta[None, :])
    # F is the function WITHOUT the weight; int f f' f'' sin^s dtheta = sum w * (f f' f'') with
    # the (1-x^2)^{(s-1)/2} factor carried by the Gauss-Jacobi weight.
    return F, w


def _ipow(k: np.ndarray) -> np.ndarray:
    return np.asarray([1.0, 1j, -1.0, -1j])[np.mod(k, 4)]


@functools.lru_cache(maxsize=8)
def coupling_matrix(btype: str, n_end: int):
    """Sparse (H*H, H2) complex matrix M with (S|R)[h', h] = sum_h'' M[(h',h), h''] * S_h''(t).

    S_h''(t) = h_{n''}(k|t|) Y_h''(t^) over harmonics of n_end2 = 2*n_end-1.
    M[(h',h),h''] = c_d * i^{n+n''-n'} * int Y_h' conj(Y_h) conj(Y_h'').
    """
    s = len(btype)
    d = s + 1
    L = n_end
    L2 = 2 * n_end - 1
    tab = index_tables(btype, L)
    tab2 = index_tables(btype, L2)
    H = tab.shape[0]
    H2 = tab2.shape[0]
    lookup2 = {tuple(r): i for i, r in enumerate(tab2.tolist())}
    cd = (2.0 * math.pi) ** (d / 2.0) * math.sqrt(2.0 / math.pi)
    inv_s2pi = 1.0 / math.sqrt(2.0 * math.pi)
    rows, cols, vals = [], [], []

    if d == 2:
        m = tab[:, 0]
        mp = m[:, None]  # h' (row)
        mm = m[None, :]  # h (col)
        m2 = mp - mm
        idx = np.vectorize(lambda v: lookup2[(v,)])(m2)
        coef = cd * _ipow(np.abs(mm) + np.abs(m2) - np.abs(mp)) * inv_s2pi
        rows = np.arange(H * H)
        M = scipy.sparse.csr_matrix((coef.ravel(), (rows, idx.ravel())), shape=(H * H, H2))
        return M

    # node tables: node i has s-1-i descendants
    Fs = {}
    for i in range(s - 1):
        desc = s - 1 - i
        Fs[desc] = _node_triple_tables(desc, L)

    @functools.lru_cache(maxsize=None)
    def triple(desc: int, lp: int, l: int, l2: int):
        """I[n', n, n''] for fixed lower indices (lp, l, l2) at a node with `desc` descendants."""
        F, w = Fs[desc]
        return np.einsum("aq,bq,cq,q->abc", F[:L, lp, :], F[:L, l, :], F[:, l2, :], w, optimize=True)

    if d == 3:
        for hp in range(H):
            n_p, m_p = tab[hp]
            for h in range(H):
                n_, m_ = tab[h]
                m2 = m_p - m_
                I = triple(1, abs(m_p), abs(m_), abs(m2))
                for n2 in range(abs(n_p - n_), n_p + n_ + 1, 2):
                    if n2 < abs(m2):
                        continue
                    g = I[n_p, n_, n2] * inv_s2pi
                    rows.append(hp * H + h)
                    cols.append(lookup2[(n2, m2)])
                    vals.append(cd * _ipow(np.asarray(n_ + n2 - n_p)) * g)
    elif d == 4:
        for hp in range(H):
            n_p, l_p, m_p = tab[hp]
            for h in range(H):
                n_, l_, m_ = tab[h]
                m2 = m_p - m_
                I1 = triple(1, abs(m_p), abs(m_), abs(m2))
                for l2 in range(abs(l_p - l_), l_p + l_ + 1, 2):
                    if l2 < abs(m2):
                        continue
                    g1 = I1[l_p, l_, l2]
                    I2 = triple(2, l_p, l_, l2)
                    for n2 in range(abs(n_p - n_), n_p + n_ + 1, 2):
                        if n2 < l2:
                            continue
                        g = I2[n_p, n_, n2] * g1 * inv_s2pi
                        rows.append(hp * H + h)
                        cols.append(lookup2[(n2, l2, m2)])
                        vals.append(cd * _ipow(np.asarray(n_ + n2 - n_p)) * g)
    else:
        # chains of any depth (d >= 5): product of one Gegenbauer triple integral per b-node; the lower indices of node i
        # are the degrees of node i + 1 (|m| for the innermost b-node).  Enumerated from the azimuth outwards.
        for hp in range(H):
            ip = tab[hp]
            for h in range(H):
                ih = tab[h]
                m2 = int(ip[s - 1] - ih[s - 1])

                def rec(i, low_p, low_h, low_2, tail, g):
                    # node i (desc = s - 1 - i), lower indices known; choose n''_i
                    n_p, n_ = int(ip[i]), int(ih[i])
                    I = triple(s - 1 - i, low_p, low_h, low_2)
                    for n2 in range(abs(n_p - n_), n_p + n_ + 1, 2):
                        if n2 < low_2:
                            continue
                        gg = g * I[n_p, n_, n2]
                        if i == 0:
                            rows.append(hp * H + h)
                            cols.append(lookup2[(n2,) + tail])
                            vals.append(cd * _ipow(np.asarray(n_ + n2 - n_p)) * gg * inv_s2pi)
                        else:
                            rec(i - 1, n_p, n_, n2, (n2,) + tail, gg)

                rec(s - 2, abs(int(ip[s - 1])), abs(int(ih[s - 1])), abs(m2), (m2,), 1.0)
    M = scipy.sparse.csr_matrix(
        (np.asarray(vals, dtype=np.complex128), (np.asarray(rows), np.asarray(cols))), shape=(H * H, H2)
    )
    return M


def translation_coef(c, t: np.ndarray, k, n_end: int) -> np.ndarray:
    """(S|R)_{h',h}(t) for translation vectors t [d, ...] -> [..., H', H]   (_biem.py:697-706)."""
    btype = _btype(c)
    d = len(btype) + 1
    L2 = 2 * n_end - 1
    sph = chain_from_cartesian([t[i] for i in range(d)])  # t is given in the chain frame
    shape = sph["r"].shape
    r = sph["r"].ravel()
    ang = [sph[i].ravel() for i in range(d - 1)]
    Y2 = harmonics(btype, ang, L2)  # [P, H2]
    hn = radial(d, L2 - 1, k * r, "h")  # [L2, P]
    deg2 = degree_table(btype, L2)
    S = hn[deg2, :].T * Y2  # [P, H2]
    M = coupling_matrix(btype, n_end)
    H = harm_count(d, n_end)
    T = (M @ S.T).T  # [P, H*H]
    return T.reshape(shape + (H, H))


# --------------------------------------------------------------------------------------
# incident fields   (_biem.py:329-450)
# --------------------------------------------------------------------------------------


def plane_wave(*, k, direction):
    k = np.asarray(k)
    direction = np.asarray(direction, dtype=np.float64)
    direction = direction / np.linalg.norm(direction, axis=0, keepdims=True)

    def inner(x):
        dd = direction[(slice(None),) + (None,) * (x.ndim - direction.ndim)]
        return np.exp(1j * k * np.sum(dd * x, axis=0))

    def inner_grad(x):
        dd = direction[(slice(None),) + (None,) * (x.ndim - direction.ndim)]
        return 1j * k * dd * np.exp(1j * k * np.sum(dd * x, axis=0))[None, ...]

    return inner, inner_grad


def point_source(*, k, source, n: int):
    k = np.asarray(k)
    source = np.asarray(source, dtype=np.float64)

    def inner(x):
        xx = x - source[(slice(None),) + (None,) * (x.ndim - source.ndim)]
        return radial(x.shape[0], n, k * np.linalg.norm(xx, axis=0), "h")[n]

    def inner_grad(x):
        xx = x - source[(slice(None),) + (None,) * (x.ndim - source.ndim)]
        r = np.linalg.norm(xx, axis=0)
        coeff = k * radial(x.shape[0], n, k * r, "h", derivative=True)[n] / r
        return coeff[None, ...] * xx

    return inner, inner_grad


# --------------------------------------------------------------------------------------
# biem / biem_u   (_biem.py:453-977) -- scalar k (the only case the reference supports with uin)
# --------------------------------------------------------------------------------------


@dataclass
class OracleResult:
    c: OracleCoordinates
    centers: np.ndarray  # [d, B]  (transposed, as the reference stores it: _biem.py:588,810)
    radii: np.ndarray
    k: complex
    n_end: int
    eta: float
    kind: str
    density: np.ndarray | None  # [B, H]
    matrix: np.ndarray | None  # [B, H, B', H']
    uin: Callable | None = None

    def uscat(self, x, far_field: bool = False, per_ball: bool = False):
        return biem_u(self, x, far_field=far_field, per_ball=per_ball)


def sd_coef(d: int, n_end: int, k, eta, radii: np.ndarray) -> np.ndarray:
    """SD_n(rho) = D - i eta S  -> [B, n_end]   (_biem.py:723-743; code, not docstring: A.7-2)."""
    jn = radial(d, n_end - 1, k * radii, "j")  # [n, B]
    jd = radial(d, n_end - 1, k * radii, "j", derivative=True)
    S = 1j * k ** (d - 2) * radii ** (d - 1) * jn
    D = 1j * k ** (d - 1) * radii ** (d - 1) * jd
    return (D - 1j * eta * S).T


def boundary_data(c, centers, radii, n_end, alpha, beta, uin, uin_grad):
    """g[q, b] at the quadrature directions   (_biem.py:611-624)."""
    co = _coords(c)
    yhat, _, _ = quadrature_nodes(co, n_end)  # chain frame
    yhat = yhat[list(co.inverse)]              # caller's cartesian frame, in which `centers` and the callables live
    x = radii[None, None, :] * yhat[:, :, None] + centers.T[:, None, :]  # [d, Q, B]
    g = np.zeros(x.shape[1:], dtype=np.complex128)
    if uin is not None:
        g = g - alpha[None, :] * uin(x)
    if uin_grad is not None:
        g = g - beta[None, :] * np.sum(uin_grad(x) * yhat[:, :, None], axis=0)
    return g


def assemble(c, centers, radii, k, n_end, eta, alpha, beta) -> np.ndarray:
    """A[b, h, b', h']   (_biem.py:692-792)."""
    btype = _btype(c)
    d = len(btype) + 1
    B = radii.shape[0]
    deg = degree_table(btype, n_end)
    H = deg.shape[0]
    SD = sd_coef(d, n_end, k, eta, radii)[:, deg]  # [B', H']
    jn = radial(d, n_end - 1, k * radii, "j").T[:, deg]  # [B, H]
    jd = radial(d, n_end - 1, k * radii, "j", derivative=True).T[:, deg]
    hn = radial(d, n_end - 1, k * radii, "h").T[:, deg]
    hd = radial(d, n_end - 1, k * radii, "h", derivative=True).T[:, deg]
    row_reg = alpha[:, None] * jn + beta[:, None] * k * jd
    row_sing = alpha[:, None] * hn + beta[:, None] * k * hd
    A = np.zeros((B, H, B, H), dtype=np.complex128)
    if B > 1:
        bi, bj = np.nonzero(~np.eye(B, dtype=bool))
        t = (centers[bi] - centers[bj]).T[list(_coords(c).axes)]  # [d, P]  t = c_b - c_b' in the chain frame
        T = translation_coef(btype, t, k, n_end)  # [P, H', H]
        A[bi, :, bj, :] = np.swapaxes(T, -1, -2) * row_reg[bi][:, :, None] * SD[bj][:, None, :]
    for b in range(B):
        A[b, np.arange(H), b, np.arange(H)] = SD[b] * row_sing[b]
    return A


def biem(
    c,
    *,
    centers,
    radii,
    k,
    n_end: int,
    alpha=1.0,
    beta=0.0,
    uin=None,
    uin_grad=None,
    eta=None,
    kind: str = "outer",
    force_matrix: bool = False,
) -> OracleResult:
    coords = _coords(c)
    btype = coords.chain
    d = coords.c_ndim
    centers = np.asarray(centers, dtype=np.float64)
    radii = np.asarray(radii, dtype=np.float64)
    if centers.shape[-1] != d:
        raise ValueError(f"The last dimension of centers must be {d}")
    k = complex(k) if np.iscomplexobj(k) else float(k)
    eta = 1.0 if eta is None else float(eta)
    B = radii.shape[0]
    alpha = np.broadcast_to(np.asarray(alpha, dtype=np.complex128), (B,))
    beta = np.broadcast_to(np.asarray(beta, dtype=np.complex128), (B,))
    deg = degree_table(btype, n_end)

    f_hat = None
    if uin is not None or uin_grad is not None:
        if np.any(alpha != 0) and uin is None:
            raise ValueError("alpha is not zero, but uin is None.")
        if np.any(beta != 0) and uin_grad is None:
            raise ValueError("beta is not zero, but uin_grad is None.")
        g = boundary_data(coords, centers, radii, n_end, alpha, beta, uin, uin_grad)
        f_hat = expand(coords, g, n_end)  # [B, H]

    use_matrix = (uin is None and uin_grad is None) or B > 1 or force_matrix
    if not use_matrix:
        SD = sd_coef(d, n_end, k, eta, radii)[:, deg]
        hn = radial(d, n_end - 1, k * radii, "h").T[:, deg]
        hd = radial(d, n_end - 1, k * radii, "h", derivative=True).T[:, deg]
        SD = SD * (alpha[:, None] * hn + beta[:, None] * hd * k)
        density = None if f_hat is None else f_hat / SD
        matrix = None
    else:
        matrix = assemble(coords, centers, radii, k, n_end, eta, alpha, beta)
        H = deg.shape[0]
        density = None
        if f_hat is not None:
            density = np.linalg.solve(matrix.reshape(B * H, B * H), f_hat.reshape(B * H)).reshape(B, H)
    return OracleResult(
        c=coords, centers=centers.T.copy(), radii=radii, k=k, n_end=n_end, eta=eta, kind=kind,
        density=density, matrix=matrix, uin=uin,
    )


def biem_u(res: OracleResult, x, far_field: bool = False, per_ball: bool = False, chunk: int = 4096):
    """Scattered field at x [d, ...(x)] -> [...(x)] (or [...(x), B])   (_biem.py:822-977)."""
    if res.density is None:
        raise ValueError("The BIEMResult does not have density.")
    btype = _btype(res.c)
    d = len(btype) + 1
    x = np.stack([np.asarray(x[i], dtype=np.float64) for i in range(d)], axis=0)
    xshape = x.shape[1:]
    xf = x.reshape(d, -1)
    P = xf.shape[1]
    B = res.radii.shape[0]
    n_end = res.n_end
    deg = degree_table(btype, n_end)
    k, eta = res.k, res.eta
    SD = sd_coef(d, n_end, k, eta, res.radii)[:, deg]  # [B, H]
    coef = res.density * SD  # [B, H]
    out = np.empty((P, B), dtype=np.complex128)
    bad = np.zeros(P, dtype=bool)
    for s0 in range(0, P, chunk):
        xs = xf[:, s0 : s0 + chunk]
        rel = xs[:, :, None] - res.centers[:, None, :]  # [d, p, B]
        sph = chain_from_cartesian([rel[a] for a in _coords(res.c).axes])  # chain frame of the tree
        r = sph["r"]
        Y = harmonics(btype, [sph[i] for i in range(d - 1)], n_end)  # [p, B, H]
        if far_field:
            rad = ((-1j) ** deg)[None, None, :]
            fac = 1.0 / (1j * k) ** ((d - 1) / 2.0) * np.exp(-1j * k * np.sum(xs[:, :, None] * res.centers[:, None, :], axis=0))
            out[s0 : s0 + chunk] = np.sum(coef[None] * rad * Y, axis=-1) * fac
        else:
            with np.errstate(all="ignore"):
                hn = radial(d, n_end - 1, k * r, "h")  # [n, p, B]
            rad = np.moveaxis(hn, 0, -1)[..., deg]
            out[s0 : s0 + chunk] = np.sum(coef[None] * rad * Y, axis=-1)
            if res.kind == "outer":
                bad[s0 : s0 + chunk] = np.any(r < res.radii[None, :], axis=-1)
            elif res.kind == "inner":
                bad[s0 : s0 + chunk] = np.any(r > res.radii[None, :], axis=-1)
            else:
                raise ValueError(f"Invalid kind: {res.kind}")
    if not per_ball:
        out = out.sum(axis=-1)
    if not far_field:
        out[bad] = np.nan
    return out.reshape(xshape + ((B,) if per_ball else ()))


def grid_centers(half: int, c_ndim: int) -> np.ndarray:
    """Synthetic geometry of the reference's sweeps (cli.py:170-185 `_center`)."""
    if half == 0:
        cen = np.zeros((2, c_ndim))
        cen[0, 1] = 2.0
        cen[1, 1] = -2.0
        return cen
    g = np.arange(-half, half) * 4 + 2
    x0, x1 = np.meshgrid(g, g, indexing="ij")
    return np.stack([x0.ravel(), x1.ravel()] + [np.zeros(x0.size)] * (c_ndim - 2), axis=-1).astype(np.float64)
