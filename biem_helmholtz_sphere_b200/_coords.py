"""Minimal stand-in for ``ultrasphere.SphericalCoordinates`` (not installable here).

The hot path only touches ``c.c_ndim``, ``c.s_ndim``, ``c.root``, ``c.from_cartesian``, ``c.to_cartesian``
and ``c.branching_types_expression_str`` (reference: _biem.py:321,593,613,617,652,699,885; plot.py:106).
Supported trees are the chains ``'a'`` (2-D), ``'ba'`` (3-D), ``'bba'`` (4-D), ... ``'bbbbbba'`` (8-D): convention decoded from the
reference's a.svg / ba.svg / bba.svg,

    x0 = r cos t0,  x1 = r sin t0 cos t1, ...,  x_{d-1} = r sin t0 ... sin t_{d-2}.

A real ultrasphere object is accepted anywhere a stand-in is: it is keyed off
``branching_types_expression_str``.
"""

from __future__ import annotations

import numpy as np

SUPPORTED = ("a", "ba", "bba", "bbba", "bbbba", "bbbbba", "bbbbbba")  # chains up to d = 8


class SphericalCoordinates:
    def __init__(self, branching_types: str):
        if branching_types not in SUPPORTED:
            raise NotImplementedError(
                f"branching types {branching_types!r}: only the chain trees {SUPPORTED} are implemented"
            )
        self.branching_types_expression_str = branching_types
        self.s_ndim = len(branching_types)
        self.c_ndim = self.s_ndim + 1
        self.root = 0

    def __repr__(self) -> str:
        return f"SphericalCoordinates({self.branching_types_expression_str!r})"

    # array-namespace agnostic (numpy arrays or torch tensors)
    @staticmethod
    def _xp(a):
        try:
            import torch

            if isinstance(a, torch.Tensor):
                return torch
        except ImportError:  # pragma: no cover
            pass
        return np

    def to_cartesian(self, spherical, as_array: bool = True):
        d = self.c_ndim
        xp = self._xp(spherical[0])
        r = spherical.get("r", 1.0)
        out = []
        prod = r
        for i in range(d - 1):
            out.append(prod * xp.cos(spherical[i]))
            prod = prod * xp.sin(spherical[i])
        out.append(prod)
        if not as_array:
            return dict(enumerate(out))
        if xp is np:
            return np.stack(np.broadcast_arrays(*out), axis=0)
        return xp.stack(xp.broadcast_tensors(*out), dim=0)

    def from_cartesian(self, x):
        d = self.c_ndim
        xp = self._xp(x[0] if not hasattr(x, "shape") else x)
        xs = [x[i] for i in range(d)]
        tails = [None] * d
        tail = xs[d - 1] * 0
        for i in range(d - 1, -1, -1):
            tail = tail + xs[i] ** 2
            tails[i] = tail
        out = {"r": xp.sqrt(tails[0])}
        at2 = np.arctan2 if xp is np else xp.atan2
        for i in range(d - 2):
            out[i] = at2(xp.sqrt(tails[i + 1]), xs[i])
        out[d - 2] = at2(xs[d - 1], xs[d - 2])
        return out


def create_from_branching_types(branching_types: str) -> SphericalCoordinates:
    """Equivalent of ``ultrasphere.create_from_branching_types`` for the supported chain trees."""
    return SphericalCoordinates(branching_types)


def branching_types_of(c) -> str:
    if isinstance(c, str):
        bt = c
    else:
        bt = getattr(c, "branching_types_expression_str", None)
    if bt is None:
        raise ValueError("c must be a SphericalCoordinates with branching_types_expression_str")
    if bt not in SUPPORTED:
        raise NotImplementedError(f"branching types {bt!r}: only the chain trees {SUPPORTED} are implemented")
    return bt
