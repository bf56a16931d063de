"""Minimal stand-in for ``ultrasphere.SphericalCoordinates`` (not installable here).

The hot path only touches ``c.c_ndim``, ``c.s_ndim``, ``c.root``, ``c.from_cartesian``, ``c.to_cartesian``
and ``c.branching_types_expression_str`` (reference: _biem.py:321,593,613,617,652,699,885; plot.py:106).

Supported trees (every tree the reference's own sweeps use, cli.py:41 ``a,ba,bpa,bba,bpbpa,caa``):

* chains of type-b / type-b' nodes over one type-a node: ``'a'`` (2-D), ``'ba'``, ``'bpa'`` (3-D), ``'bba'``, ``'bpbpa'``,
  ``'bbpa'``, ... up to 8-D.  A type-b node sends cos to its leaf and sin to the sub-tree, theta in [0, pi]; a type-b'
  node (written ``bp``) sends SIN to its leaf and cos to the sub-tree, theta in [-pi/2, pi/2]:

      'ba' : x0 = r cos t0,  x1 = r sin t0 cos t1,  x2 = r sin t0 sin t1
      'bpa': x0 = r sin t0,  x1 = r cos t0 cos t1,  x2 = r cos t0 sin t1

* ``'caa'`` (4-D, Hopf): x0 = r cos t0 cos t1, x1 = r cos t0 sin t1, x2 = r sin t0 cos t2, x3 = r sin t0 sin t2,
  t0 in [0, pi/2].

Conventions decoded from the reference's a.svg / ba.svg / bpa.svg / bba.svg / bpbpa.svg / caa.svg and pinned by a
brute-force search against the golden rows of jascome/jascome_output.csv (tools/tree_convention_search.py).

``relabel({0: d - 1, d - 1: 0})`` renames cartesian leaves the way the reference CLI does for the ``p`` trees
(cli.py:63-69, ``nx.relabel_nodes``): it is the CLI's relabelled ``bpa`` whose polar axis is x2 that the golden rows hold.

What a tree changes numerically.  All trees of one dimension span the SAME harmonic space (degree < n_end on S^{d-1}),
and every result the reference exposes through ``uscat`` is basis invariant; the tree decides (i) which cartesian axis
plays which role -- b / b' chains are the plain chain in a permuted frame (``axes``) -- and (ii) where ``ush.expand``
samples the boundary data (``tree``: the chain product rule or the Hopf rule of 'caa').  The device code therefore works
in the chain frame throughout; densities are expressed in the chain basis of that frame (harmonic ordering is not pinned
by any reference artefact).

A real ultrasphere object is accepted anywhere a stand-in is: it is keyed off ``branching_types_expression_str``
(an axis relabelling of such an object cannot be seen and is taken to be the identity).
"""

from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np

MAX_C_NDIM = 8
TREE_CHAIN, TREE_HOPF = 0, 1
_TOKEN = re.compile(r"bp|b'|a|b|c")


def _tokens(branching_types: str) -> list[str]:
    pos, out = 0, []
    while pos < len(branching_types):
        m = _TOKEN.match(branching_types, pos)
        if not m:
            raise ValueError(f"invalid branching types {branching_types!r}")
        out.append("bp" if m.group(0) in ("bp", "b'") else m.group(0))
        pos = m.end()
    return out


@dataclass(frozen=True)
class TreeSpec:
    """What the hot path needs to know about a coordinate tree."""

    d: int                   # cartesian dimension
    chain: str               # chain tree of the same dimension whose basis the device code uses ('a', 'ba', 'bba', ...)
    axes: tuple[int, ...]    # chain coordinate i is the caller's cartesian coordinate axes[i]
    tree: int                # TREE_CHAIN / TREE_HOPF: right-hand-side quadrature rule
    nodes: tuple[str, ...]   # node types in tree order

    @property
    def identity(self) -> bool:
        return self.axes == tuple(range(self.d))

    @property
    def inverse(self) -> tuple[int, ...]:
        return tuple(int(i) for i in np.argsort(self.axes))


def _parse(branching_types: str) -> tuple[tuple[str, ...], int]:
    toks = tuple(_tokens(branching_types))
    if toks == ("c", "a", "a"):
        return toks, TREE_HOPF
    if toks and toks[-1] == "a" and all(t in ("b", "bp") for t in toks[:-1]) and len(toks) + 1 <= MAX_C_NDIM:
        return toks, TREE_CHAIN
    raise NotImplementedError(
        f"branching types {branching_types!r}: implemented are chains of b / bp nodes over one a node (up to "
        f"{MAX_C_NDIM}-D: 'a', 'ba', 'bpa', 'bba', 'bpbpa', ...) and 'caa'"
    )


class SphericalCoordinates:
    def __init__(self, branching_types: str, axes: tuple[int, ...] | None = None):
        self.nodes, self.tree = _parse(branching_types)
        self.branching_types_expression_str = branching_types
        self.s_ndim = len(self.nodes)
        self.c_ndim = self.s_ndim + 1
        self.root = 0
        axes = tuple(range(self.c_ndim)) if axes is None else tuple(int(a) for a in axes)
        if sorted(axes) != list(range(self.c_ndim)):
            raise ValueError(f"axes must be a permutation of 0..{self.c_ndim - 1}, got {axes}")
        self.axes = axes

    def __repr__(self) -> str:
        extra = "" if self.axes == tuple(range(self.c_ndim)) else f", axes={self.axes}"
        return f"SphericalCoordinates({self.branching_types_expression_str!r}{extra})"

    def relabel(self, mapping: dict[int, int]) -> "SphericalCoordinates":
        """Rename cartesian leaves (``nx.relabel_nodes`` of the reference CLI, cli.py:63-69)."""
        return SphericalCoordinates(self.branching_types_expression_str, tuple(mapping.get(a, a) for a in self.axes))

    @property
    def spec(self) -> TreeSpec:
        chain = "b" * (self.c_ndim - 2) + "a"
        return TreeSpec(d=self.c_ndim, chain=chain, axes=self.axes, tree=self.tree, nodes=self.nodes)

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
        y = []
        if self.tree == TREE_HOPF:
            c0, s0 = xp.cos(spherical[0]), xp.sin(spherical[0])
            y = [r * c0 * xp.cos(spherical[1]), r * c0 * xp.sin(spherical[1]),
                 r * s0 * xp.cos(spherical[2]), r * s0 * xp.sin(spherical[2])]
        else:
            prod = r
            for i in range(d - 2):
                if self.nodes[i] == "b":
                    y.append(prod * xp.cos(spherical[i]))
                    prod = prod * xp.sin(spherical[i])
                else:
                    y.append(prod * xp.sin(spherical[i]))
                    prod = prod * xp.cos(spherical[i])
            y.append(prod * xp.cos(spherical[d - 2]))
            y.append(prod * xp.sin(spherical[d - 2]))
        out = [None] * d
        for i in range(d):
            out[self.axes[i]] = y[i]
        if not as_array:
            return dict(enumerate(out))
        if xp is np:
            return np.stack(np.broadcast_arrays(*out), axis=0)
        return xp.stack(xp.broadcast_tensors(*out), dim=0)

    def from_cartesian(self, x):
        d = self.c_ndim
        xp = self._xp(x[0] if not hasattr(x, "shape") else x)
        ys = [x[self.axes[i]] for i in range(d)]
        at2 = np.arctan2 if xp is np else xp.atan2
        if self.tree == TREE_HOPF:
            ra = xp.sqrt(ys[0] ** 2 + ys[1] ** 2)
            rb = xp.sqrt(ys[2] ** 2 + ys[3] ** 2)
            return {"r": xp.sqrt(ra**2 + rb**2), 0: at2(rb, ra), 1: at2(ys[1], ys[0]), 2: at2(ys[3], ys[2])}
        tails = [None] * d
        tail = ys[d - 1] * 0
        for i in range(d - 1, -1, -1):
            tail = tail + ys[i] ** 2
            tails[i] = tail
        out = {"r": xp.sqrt(tails[0])}
        for i in range(d - 2):
            if self.nodes[i] == "b":
                out[i] = at2(xp.sqrt(tails[i + 1]), ys[i])
            else:
                out[i] = at2(ys[i], xp.sqrt(tails[i + 1]))
        out[d - 2] = at2(ys[d - 1], ys[d - 2])
        return out


def create_from_branching_types(branching_types: str) -> SphericalCoordinates:
    """Equivalent of ``ultrasphere.create_from_branching_types`` for the supported trees."""
    return SphericalCoordinates(branching_types)


def tree_spec(c) -> TreeSpec:
    """TreeSpec of a stand-in, of a branching-types string, or of a real ultrasphere object."""
    if isinstance(c, SphericalCoordinates):
        return c.spec
    bt = c if isinstance(c, str) else getattr(c, "branching_types_expression_str", None)
    if bt is None:
        raise ValueError("c must be a SphericalCoordinates with branching_types_expression_str")
    return SphericalCoordinates(bt).spec


def branching_types_of(c) -> str:
    """Chain tree ('a', 'ba', 'bba', ...) whose basis the device code uses for ``c``: len(.) + 1 == c_ndim."""
    return tree_spec(c).chain
