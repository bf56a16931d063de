"""Chains beyond 4-D ('bbba' = 5-D, 'bbbba' = 6-D): the generic coupling enumeration (one Gegenbauer triple integral per
b-node), the hyperspherical radial functions and the generic harmonics / field kernels against the oracle, whose own
5-D / 6-D coupling is validated from first principles (addition theorem, tests/test_oracle_identities.py)."""
import numpy as np
import pytest

from oracle import biem_oracle as bo

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("btype,n_end,B", [("bbba", 4, 2), ("bbba", 5, 3), ("bbbba", 3, 2)])
def test_high_dimensional_chain(btype, n_end, B):
    import biem_helmholtz_sphere_b200 as bhs
    from biem_helmholtz_sphere_b200 import _ops

    c = bhs.create_from_branching_types(btype)
    d = c.c_ndim
    # index tables bit-exact, harmonics to round-off
    plan = _ops.get_plan(d, n_end)
    assert np.array_equal(plan.index_table(), bo.index_tables(btype, n_end).astype(np.int32))
    rng = np.random.default_rng(d)
    x = rng.normal(size=(d, 40))
    sph = bo.OracleCoordinates(btype).from_cartesian(x)
    Yo = bo.harmonics(btype, [sph[i] for i in range(d - 1)], n_end)
    Yg = _ops.harmonics(d, n_end, x).cpu().numpy()
    assert np.max(np.abs(Yg - Yo)) < 1e-13 * np.max(np.abs(Yo))
    # full path: rhs + assembly + solve + field
    cen = np.zeros((B, d))
    cen[:, 1] = 4.0 * np.arange(B) - 2.0 * (B - 1)
    cen[:, 3] = 0.3 * np.arange(B)
    rad = np.linspace(1.0, 0.8, B)
    k = 1.2
    dirn = np.zeros(d)
    dirn[0], dirn[2] = 0.6, 0.8
    uin, ug = bhs.plane_wave(k=np.asarray(k), direction=dirn)
    calc = bhs.biem(c, uin=uin, uin_grad=ug, k=np.asarray(k), n_end=n_end, eta=np.asarray(0.9), centers=cen, radii=rad,
                    alpha=1.0, beta=0.3)
    ou, og = bo.plane_wave(k=k, direction=dirn)
    ref = bo.biem(btype, uin=ou, uin_grad=og, k=k, n_end=n_end, eta=0.9, centers=cen, radii=rad, alpha=1.0, beta=0.3)
    assert rel(calc.matrix.reshape(ref.matrix.shape), ref.matrix) < TOL
    assert rel(calc.density, ref.density) < TOL
    xp = rng.uniform(-5, 5, size=(d, 30))
    u, want = calc.uscat(xp), ref.uscat(xp)
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(u), ~ok) and rel(u[ok], want[ok]) < TOL
    xh = xp / np.linalg.norm(xp, axis=0, keepdims=True)
    assert rel(calc.uscat(xh, far_field=True), ref.uscat(xh, far_field=True)) < TOL
