"""How the conventions of the b' / c coordinate trees were recovered.

The reference's `ultrasphere` package is not installable here, but its CLI left golden rows for the trees `bpa`, `bpbpa`
and `caa` (jascome/jascome_output.csv:2-12,19-27; cli.py:41,63-69).  Every tree of one dimension spans the same harmonic
space, so the rows can only depend on (i) which cartesian axis plays which role and (ii) where the right-hand-side
quadrature samples the sphere.  Both sets are small and discrete; this script enumerates them with the CPU oracle and
prints the relative error against the golden rows with n_end = 1..4 (n_end >= 2 discriminates at 1e-3 .. 1e-6).

Result (oracle/biem_oracle.py, biem_helmholtz_sphere_b200/_coords.py implement it):
  bpa   = chain 'ba'  in the frame x_chain = x[(2, 1, 0)]        (b' leaf numbered like b's; the CLI swaps axes 0 <-> d-1)
  bpbpa = chain 'bba' in the frame x_chain = x[(3, 1, 2, 0)]
  caa   = chain-'bba' harmonic space, right-hand side sampled by the Hopf product rule: n_end Gauss-Legendre nodes in
          cos(2 theta_0), 2 n_end equispaced nodes on each circle, frame (0, 1, 2, 3)
all to <= 2e-15 for n_end <= 4.

    python tools/tree_convention_search.py          (CPU only, ~1 min)
"""
import itertools
import math
import os
import sys

import numpy as np
from scipy import special as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import load  # noqa: E402
from oracle import biem_oracle as O  # noqa: E402

rows = load("jascome_output.csv")


def gold(bt):
    return {r["n_end"]: r["uscat"] for r in rows if r["branching_types"] == bt}


def run_chain_in_frame(base, n_end, perm):
    d = len(base) + 1
    dirv = np.array([1.0] + [0.0] * (d - 1))
    cen = O.grid_centers(0, d)
    uin, _ = O.plane_wave(k=1.0, direction=dirv[list(perm)])
    res = O.biem(base, uin=uin, k=1.0, n_end=n_end, eta=1.0, centers=cen[:, list(perm)], radii=np.ones(2))
    return complex(res.uscat(np.zeros(d)))


def hopf_rule(n, rule):
    if rule == "gauss-legendre in cos(2 t0)":
        t, w = sp.roots_legendre(n); th0 = 0.5 * np.arccos(t); w0 = w / 4.0
    elif rule == "gauss-jacobi(0,1) in cos(t0)":
        t, w = sp.roots_jacobi(n, 0.0, 1.0); th0 = np.arccos((t + 1) / 2); w0 = w / 4.0
    elif rule == "gauss-jacobi(0,1) in sin(t0)":
        t, w = sp.roots_jacobi(n, 0.0, 1.0); th0 = np.arcsin((t + 1) / 2); w0 = w / 4.0
    elif rule == "gauss-legendre in cos(t0)":
        x, w = sp.roots_legendre(n); x = (x + 1) / 2; th0 = np.arccos(x); w0 = w / 2 * x
    else:
        x, w = sp.roots_legendre(n); x = (x + 1) / 2; th0 = np.arcsin(x); w0 = w / 2 * x
    az = 2 * math.pi * np.arange(2 * n) / (2 * n)
    T0, T1, T2 = (a.ravel() for a in np.meshgrid(th0, az, az, indexing="ij"))
    W = (w0[:, None, None] * np.full((1, 2 * n, 1), math.pi / n) * np.full((1, 1, 2 * n), math.pi / n)).ravel()
    y = np.stack([np.cos(T0) * np.cos(T1), np.cos(T0) * np.sin(T1), np.sin(T0) * np.cos(T2), np.sin(T0) * np.sin(T2)])
    return y, W


def run_hopf(n_end, rule, perm):
    d, bt = 4, "bba"
    y, W = hopf_rule(n_end, rule)
    yhat = y[list(perm)]
    cen, rad = O.grid_centers(0, d), np.ones(2)
    uin, _ = O.plane_wave(k=1.0, direction=np.array([1.0, 0, 0, 0]))
    gv = -uin(rad[None, None, :] * yhat[:, :, None] + cen.T[:, None, :])
    sph = O.chain_from_cartesian(list(yhat))
    Y = O.harmonics(bt, [sph[i] for i in range(d - 1)], n_end)
    f_hat = np.einsum("q,qb,qh->bh", W, gv, np.conj(Y))
    A = O.assemble(bt, cen, rad, 1.0, n_end, 1.0, np.ones(2, complex), np.zeros(2, complex))
    H = Y.shape[1]
    dens = np.linalg.solve(A.reshape(2 * H, 2 * H), f_hat.reshape(2 * H)).reshape(2, H)
    res = O.OracleResult(c=O.OracleCoordinates(bt), centers=cen.T.copy(), radii=rad, k=1.0, n_end=n_end, eta=1.0,
                         kind="outer", density=dens, matrix=None)
    return complex(res.uscat(np.zeros(d)))


if __name__ == "__main__":
    for bt, base in (("bpa", "ba"), ("bpbpa", "bba")):
        g = gold(bt)
        for perm in itertools.permutations(range(len(base) + 1)):
            errs = [abs(run_chain_in_frame(base, n, perm) - g[n]) / abs(g[n]) for n in (1, 2, 3, 4)]
            print(f"{bt:6s} chain {base:4s} frame {perm}: " + " ".join(f"{e:.1e}" for e in errs) + ("   <== match" if max(errs) < 1e-13 else ""), flush=True)
    g = gold("caa")
    for rule in ("gauss-legendre in cos(2 t0)", "gauss-jacobi(0,1) in cos(t0)", "gauss-jacobi(0,1) in sin(t0)",
                 "gauss-legendre in cos(t0)", "gauss-legendre in sin(t0)"):
        for perm in ((0, 1, 2, 3), (2, 3, 0, 1), (0, 2, 1, 3), (1, 3, 0, 2), (0, 3, 1, 2), (1, 2, 0, 3)):
            errs = [abs(run_hopf(n, rule, perm) - g[n]) / abs(g[n]) for n in (1, 2, 3, 4)]
            print(f"caa    {rule:30s} frame {perm}: " + " ".join(f"{e:.1e}" for e in errs) + ("   <== match" if max(errs) < 1e-13 else ""), flush=True)
