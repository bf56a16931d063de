"""Host-side logic that needs no GPU: synthetic geometries, the coordinate stand-in, the memory model, the input checks
of the API mirror (they run before anything touches the device) and the no-CPU-fallback guarantee."""
import numpy as np
import pytest

import biem_helmholtz_sphere_b200 as bhs
from biem_helmholtz_sphere_b200 import _biem, geometry
from oracle import biem_oracle as bo


def test_grid_centers_match_reference_center_function():
    # cli._center (cli.py:170-185) restated in the oracle; the package's copy must agree exactly
    for half in (0, 1, 2, 4):
        for d in (2, 3, 4):
            assert np.array_equal(geometry.grid_centers(half, d), bo.grid_centers(half, d))
    assert geometry.grid_centers(2, 3).shape == (16, 3)
    ks = geometry.sweep_wavenumbers(256)
    assert ks[0] == 0.5 and ks[-1] == 8.0 and len(ks) == 256
    x = geometry.probe_ring(64, 10.0, 3)
    assert x.shape == (3, 65) and np.allclose(np.linalg.norm(x[:, 1:], axis=0), 10.0) and not x[:, 0].any()
    g = geometry.field_grid(8, 20.0, 3)
    assert g.shape == (3, 8, 8) and g[0, 0, 0] == -20.0 and g[1, -1, -1] == 20.0 and not g[2].any()


@pytest.mark.parametrize("bt", ["a", "ba", "bba", "bpa", "bpbpa", "bbpa", "caa"])
@pytest.mark.parametrize("relabel", [False, True])
def test_coordinate_stand_in_round_trip(bt, relabel):
    """Every tree of the reference's sweeps (cli.py:41), with and without the CLI's 0 <-> d-1 leaf relabelling
    (cli.py:63-69): cartesian -> angles -> cartesian, angle ranges, and agreement with the oracle's coordinates."""
    c = bhs.create_from_branching_types(bt)
    o = bo.OracleCoordinates(bt)
    d = c.c_ndim
    if relabel:
        c, o = c.relabel({0: d - 1, d - 1: 0}), o.relabel({0: d - 1, d - 1: 0})
    ntok = len(bt.replace("bp", "b"))
    assert c.c_ndim == ntok + 1 and c.s_ndim == ntok and c.branching_types_expression_str == bt
    rng = np.random.default_rng(0)
    x = rng.normal(size=(d, 50))
    sph = c.from_cartesian(x)
    back = c.to_cartesian(sph)
    assert np.allclose(back, x, atol=1e-13)
    oo = o.from_cartesian(x)
    for k in sph:
        assert np.allclose(sph[k], oo[k])
    assert np.allclose(o.to_cartesian(oo), x, atol=1e-13)
    for i, node in enumerate(c.nodes[:-1] if bt != "caa" else c.nodes[:1]):
        lo, hi = {"b": (0.0, np.pi), "bp": (-np.pi / 2, np.pi / 2), "c": (0.0, np.pi / 2)}[node]
        assert np.all(sph[i] >= lo) and np.all(sph[i] <= hi)
    spec = c.spec
    assert spec.d == d and len(spec.chain) == d - 1 and sorted(spec.axes) == list(range(d))
    assert spec.tree == (1 if bt == "caa" else 0)
    # the chain frame: x_chain = x[axes] has the plain chain's angles up to theta -> pi/2 - theta at b' nodes
    if bt != "caa":
        ch = bo.chain_from_cartesian([x[a] for a in spec.axes])
        for i, node in enumerate(c.nodes[:-1]):
            assert np.allclose(sph[i], ch[i] if node == "b" else np.pi / 2 - ch[i])
        assert np.allclose(sph[d - 2], ch[d - 2])


def test_coordinate_stand_in_rejects_unknown_trees():
    for bad in ("cba", "ab", "bb", "caaa", "x"):
        with pytest.raises((NotImplementedError, ValueError)):
            bhs.create_from_branching_types(bad)
    with pytest.raises(ValueError):
        bhs.SphericalCoordinates("ba", axes=(0, 0, 1))


def test_memory_model_and_harmonic_counts():
    # harm_n_ndim_le: 2-D 2n-1, 3-D n^2, 4-D n(n+1)(2n+1)/6   
    for n in range(1, 12):
        assert _biem.harm_n_ndim_le(n, c_ndim=2) == 2 * n - 1
        assert _biem.harm_n_ndim_le(n, c_ndim=3) == n * n
        assert _biem.harm_n_ndim_le(n, c_ndim=4) == n * (n + 1) * (2 * n + 1) // 6
        assert _biem.harm_n_ndim_le(n, c_ndim=3) == bo.harm_count(3, n)
    # max_memory / max_n_end keep the reference's model (_biem.py:23-74), including the element count for d <= 3
    assert bhs.max_memory(c_ndim=3, n_end=16, n_balls=16) == 16 ** 2 * 256 ** 2
    assert bhs.max_memory(c_ndim=4, n_end=3, n_balls=2) == 4 * (5 * 27) ** 2 * (11 * 6 ** 3) * 16
    lim = bhs.max_memory(c_ndim=3, n_end=10, n_balls=4)
    assert bhs.max_n_end(c_ndim=3, memory_limit=lim, n_balls=4) == 10


def test_plane_wave_closures_and_shape_errors():
    u, g = bhs.plane_wave(k=np.asarray(2.0), direction=np.array([3.0, 0.0, 4.0]))
    x = np.array([[1.0, 0.0], [0.0, 1.0], [0.5, -0.5]])
    assert np.allclose(u(x), np.exp(2j * (0.6 * x[0] + 0.8 * x[2])))
    assert np.allclose(g(x), 2j * np.array([0.6, 0.0, 0.8])[:, None] * u(x)[None])
    with pytest.raises(ValueError):
        bhs.plane_wave(k=np.asarray(1.0), direction=np.ones((3, 2)))
    with pytest.raises(ValueError):
        bhs.point_source(k=np.asarray([1.0, 2.0]), source=np.zeros(3), n=0)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    c = bhs.create_from_branching_types("ba")
    with pytest.raises(Exception) as ei:
        bhs.biem(c, k=np.asarray(1.0), n_end=3, centers=np.array([[0.0, 2.0, 0.0], [0.0, -2.0, 0.0]]), radii=np.ones(2))
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_panel_row_map_algorithm():
    """Specification of the pivot bookkeeping of the last tournament round (csrc/lu.cu, select_finish): from the ordered list
    of pivot rows of a panel it derives, in O(w) steps, the LAPACK-style sequential swaps ipiv[j + c] and the panel's net row
    map used by lu_permute_kernel.  Restated here in Python and checked against literally applying the swaps."""
    import random

    def finish(j, w, s_win):
        where = [j + t for t in range(32)]   # current position of diagonal-block row j + t
        cont = [j + t for t in range(32)]    # row currently held by diagonal-block position j + t
        piv = []
        for c in range(w):
            r = s_win[c]
            loc = where[r - j] if r < j + w else r
            d = cont[c]
            where[d - j] = loc
            if loc < j + w:
                cont[loc - j] = d
            if r < j + w:
                where[r - j] = j + c
            piv.append(loc)
        src_top = [s_win[q] if q < w else j + q for q in range(32)]
        dst_out = [where[t] if (t < w and where[t] >= j + w) else -1 for t in range(32)]
        return piv, src_top, dst_out

    rnd = random.Random(7)
    for _ in range(3000):
        w = rnd.choice([32, 32, 17, 5, 2, 1])
        j = rnd.randrange(0, 50)
        m = w if (w < 32 or rnd.random() < 0.1) else w + rnd.choice([1, 7, 64, 400])
        s_win = rnd.sample(range(j, j + m), w)
        if rnd.random() < 0.4:  # bias towards pivots that already sit in the diagonal block
            s_win = [r if rnd.random() < 0.5 else j + q for q, r in enumerate(s_win)]
            if len(set(s_win)) < w:
                continue
        piv, src_top, dst_out = finish(j, w, s_win)
        rows = list(range(j + m))
        for c, p in enumerate(piv):  # LAPACK semantics
            rows[j + c], rows[p] = rows[p], rows[j + c]
        assert rows[j : j + w] == s_win
        net = list(range(j + m))
        for q in range(w):
            net[j + q] = src_top[q]
        for t in range(w):
            if dst_out[t] >= 0:
                net[dst_out[t]] = j + t
        assert net == rows
