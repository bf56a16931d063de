"""The two plot-shaped callers of ``uscat(per_ball=True[, far_field=True])`` -- the data side of the reference's
``plot_biem`` / ``plot_biem_far`` (src/biem_helmholtz_sphere/plot.py:12-134, :137-217) without the plotly figure.

They are the only in-tree consumers of the large-grid evaluation path: a 2-D heat map of
``Re[(u_in + sum_b u_b) e^{-2 pi i t}]`` over a coordinate plane, and the polar far-field pattern ``|sum_b u_b^inf|`` on the
unit circle of a coordinate plane.  Each function returns exactly the arrays (and the title string) the reference hands to
``px.imshow`` / ``px.line_polar`` (plot.py:117-128, :205-216), so a front-end only has to draw them.

One deliberate difference: the reference builds its grid by a round trip ``c.to_cartesian(c.from_cartesian(...))``
(plot.py:72-78), which perturbs the exact zeros of the other coordinates by a few ulps; here the grid is built directly, so
a heat map through coplanar centres keeps its points exactly in the plane and the planar field kernel serves it.
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any

import numpy as np

from ._biem import BIEMResultCalculator

__all__ = ["heatmap_field", "far_field_pattern"]


def _title_tail(biem_res: Any) -> str:
    c = biem_res.c
    k, eta = np.asarray(biem_res.k), np.asarray(biem_res.eta)
    kk = complex(k) if np.iscomplexobj(k) else float(k)
    ee = complex(eta) if np.iscomplexobj(eta) else float(eta)
    return (f"{c.c_ndim:g}D, type {c.branching_types_expression_str} coordinates, Max Degree={biem_res.n_end - 1:g}, "
            f"k={kk:g}, η={ee:g}")


def heatmap_field(
    biem_res: BIEMResultCalculator,
    /,
    *,
    plot_uin: bool = True,
    plot_uscateach: bool | Sequence[bool] = True,
    xspace: tuple[float, float, int] | None = None,
    yspace: tuple[float, float, int] | None = None,
    n_t: int = 1,
    xaxis: int = 0,
    yaxis: int = 1,
    log: bool = False,
) -> dict:
    """Arrays of the reference's ``plot_biem`` (plot.py:61-99).

    Returns ``{"x" [nx], "y" [ny], "cartesian" [c_ndim, nx, ny], "uin" [nx, ny], "uscateach" [nx, ny, B],
    "uplot_re" [n_t, nx, ny], "title"}``; ``uplot_re`` is what ``px.imshow`` receives (after its ``moveaxis``), NaN inside
    the balls (outside for ``kind="inner"``).
    """
    xspace_ = xspace or (-1, 1, 100)
    yspace_ = yspace or (-1, 1, 100)
    sel = np.asarray(plot_uscateach)
    if sel.ndim == 0:
        sel = sel[None]
    d = biem_res.c.c_ndim
    x = np.linspace(*xspace_, dtype=np.float64)
    y = np.linspace(*yspace_, dtype=np.float64)
    cart = np.zeros((d, x.size, y.size))
    cart[xaxis] = x[:, None]
    cart[yaxis] = y[None, :]
    uin = np.zeros(cart.shape[1:], dtype=np.complex128) if biem_res.uin is None else np.asarray(biem_res.uin(cart))
    uscateach = np.asarray(biem_res.uscat(cart, per_ball=True))  # [nx, ny, B]
    t = np.arange(n_t, dtype=np.float64)[:, None, None] / n_t
    texp = np.exp(-1j * t * 2 * np.pi)
    uplot = plot_uin * uin + np.sum(sel[None, None, :] * uscateach, axis=-1)
    uplot_re = np.real(uplot * texp)
    if log:
        uplot_re = np.sign(uplot_re) * np.log1p(np.abs(uplot_re))
    title = ""
    if plot_uin:
        title += "Incident Field"
    if np.any(sel):
        if plot_uin:
            title += " + "
        shown = np.nonzero(np.broadcast_to(sel, (uscateach.shape[-1],)))[0]
        title += "Scattered Field by Ball " + ", ".join(str(int(i)) for i in shown)
    title += "<br>" + _title_tail(biem_res)
    return {"x": x, "y": y, "cartesian": cart, "uin": uin, "uscateach": uscateach, "uplot_re": uplot_re, "title": title}


def far_field_pattern(
    biem_res: BIEMResultCalculator,
    /,
    *,
    plot_uscateach: bool | Sequence[bool] = True,
    n_points: int = 100,
    xaxis: int = 0,
    yaxis: int = 1,
) -> dict:
    """Arrays of the reference's ``plot_biem_far`` (plot.py:170-216): the per-ball far-field patterns at ``n_points`` unit
    directions of the (xaxis, yaxis) plane and the magnitude of their (selected) sum.

    Returns ``{"theta" [n_points] (radians), "theta_deg", "cartesian" [c_ndim, n_points], "uscateach" [n_points, B],
    "uplot_abs" [n_points], "title"}``.
    """
    sel = np.asarray(plot_uscateach)
    if sel.ndim == 0:
        sel = sel[None]
    d = biem_res.c.c_ndim
    theta = np.arange(n_points, dtype=np.float64) * (2 * np.pi / n_points)
    cart = np.zeros((d, n_points))
    cart[xaxis] = np.cos(theta)
    cart[yaxis] = np.sin(theta)
    uscateach = np.asarray(biem_res.uscat(cart, per_ball=True, far_field=True))  # [n_points, B]
    uplot_abs = np.abs(np.sum(sel[None, :] * uscateach, axis=-1))
    shown = np.nonzero(np.broadcast_to(sel, (uscateach.shape[-1],)))[0]
    title = "Far Field Pattern by Ball " + ", ".join(str(int(i)) for i in shown) + "<br>" + _title_tail(biem_res)
    return {"theta": theta, "theta_deg": theta * 180 / np.pi, "cartesian": cart, "uscateach": uscateach,
            "uplot_abs": uplot_abs, "title": title}
