"""Synthetic geometries of the reference's sweeps and of the benchmark configs.

``grid_centers`` restates ``cli._center`` (reference src/biem_helmholtz_sphere/cli.py:170-185): ``half == 0`` is the
two-sphere case (0, +-2, 0, ...); otherwise a (2 half) x (2 half) grid of centres {4 i + 2 : i = -half..half-1}^2 in
the first two axes, zeros elsewhere.  Pure NumPy, host side only.
"""

from __future__ import annotations

import numpy as np


def grid_centers(half: int, c_ndim: int) -> np.ndarray:
    if half == 0:
        cen = np.zeros((2, c_ndim))
        cen[0, 1] = 2.0
        cen[1, 1] = -2.0
        return cen
    g = np.arange(-half, half) * 4.0 + 2.0
    x0, x1 = np.meshgrid(g, g, indexing="ij")
    cols = [x0.ravel(), x1.ravel()] + [np.zeros(x0.size)] * (c_ndim - 2)
    return np.stack(cols, axis=-1)


def sweep_wavenumbers(n: int = 256, k_lo: float = 0.5, k_hi: float = 8.0) -> np.ndarray:
    """The C3 wavenumber sweep k_i = k_lo + (k_hi - k_lo) i / (n - 1)."""
    if n == 1:
        return np.asarray([k_lo])
    return k_lo + (k_hi - k_lo) * np.arange(n) / (n - 1)


def probe_ring(n: int, radius: float, c_ndim: int) -> np.ndarray:
    """Origin + n points on a circle of the given radius in the (x0, x1) plane -> [c_ndim, n + 1]."""
    th = 2.0 * np.pi * (np.arange(n) + 0.5) / n
    x = np.zeros((c_ndim, n + 1))
    x[0, 1:] = radius * np.cos(th)
    x[1, 1:] = radius * np.sin(th)
    return x


def field_grid(n: int, extent: float, c_ndim: int) -> np.ndarray:
    """n x n grid on the x2 = ... = 0 plane over [-extent, extent]^2 -> [c_ndim, n, n] (C5 heat map)."""
    g = np.linspace(-extent, extent, n)
    x0, x1 = np.meshgrid(g, g, indexing="ij")
    x = np.zeros((c_ndim, n, n))
    x[0] = x0
    x[1] = x1
    return x
