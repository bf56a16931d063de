"""Complex wavenumbers (absorbing medium, Im k > 0) in 3-D against the oracle (scipy's complex-argument Bessel
functions).  No reference golden data exists for this case: parity is oracle-pinned, tolerance 1e-10 relative."""
import numpy as np
import pytest

from oracle import biem_oracle as bo

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def bhs():
    import biem_helmholtz_sphere_b200 as m

    return m


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


CEN = np.array([[0.0, 2.0, 0.3], [0.5, -2.0, 0.0], [4.0, 0.0, -1.0]])
RAD = np.array([1.0, 0.8, 1.2])
DIRN = np.array([0.3, -0.5, 0.8])


@pytest.mark.parametrize("k", [1.3 + 0.2j, 0.4 + 1.5j, 6.0 + 0.05j, 2.5 + 3.0j])
@pytest.mark.parametrize("n_end", [6, 14])
def test_complex_k_density_and_fields(bhs, k, n_end):
    c = bhs.create_from_branching_types("ba")
    uin, uin_grad = bhs.plane_wave(k=np.asarray(k), direction=DIRN)
    calc = bhs.biem(c, uin=uin, uin_grad=uin_grad, k=np.asarray(k), n_end=n_end, eta=np.asarray(0.7), centers=CEN,
                    radii=RAD, alpha=1.0, beta=0.3 + 0.1j)
    ou, og = bo.plane_wave(k=k, direction=DIRN)
    ref = bo.biem("ba", uin=ou, uin_grad=og, k=k, n_end=n_end, eta=0.7, centers=CEN, radii=RAD, alpha=1.0, beta=0.3 + 0.1j)
    assert calc.k.dtype == np.complex128 and complex(calc.k) == k
    assert rel(calc.matrix.reshape(ref.matrix.shape), ref.matrix) < TOL
    assert rel(calc.density, ref.density) < TOL
    rng = np.random.default_rng(3)
    x = rng.uniform(-6, 6, size=(3, 60))
    x[:, 0] = 0.0
    u, want = calc.uscat(x), ref.uscat(x)
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(u), ~ok)
    assert rel(u[ok], want[ok]) < TOL
    pb, wpb = calc.uscat(x, per_ball=True), ref.uscat(x, per_ball=True)
    assert rel(pb[ok], wpb[ok]) < TOL
    xh = x[:, 1:] / np.linalg.norm(x[:, 1:], axis=0, keepdims=True)
    assert rel(calc.uscat(xh, far_field=True), ref.uscat(xh, far_field=True)) < TOL


def test_complex_k_sweep_shortcut_and_limits(bhs):
    c = bhs.create_from_branching_types("ba")
    ks = np.array([0.7 + 0.1j, 1.3 + 0.4j, 2.9 + 0.0j, 5.1 + 1.0j])
    uin = bhs.plane_wave(k=ks, direction=DIRN[:, None])[0]
    calc = bhs.biem(c, uin=uin, k=ks, n_end=9, eta=np.ones(4), centers=CEN[None], radii=RAD[None], keep_matrix=False)
    x = np.array([[0.0, 6.0, -3.0], [0.0, 0.5, 4.0], [0.0, 1.0, -2.0]])
    u = calc.uscat(x)
    for i, k in enumerate(ks):
        ref = bo.biem("ba", uin=bo.plane_wave(k=k, direction=DIRN)[0], k=k, n_end=9, eta=1.0, centers=CEN, radii=RAD)
        assert rel(calc.density[i], ref.density) < TOL
        assert rel(u[:, i], ref.uscat(x)) < TOL
    # single-sphere shortcut with a complex wavenumber
    k = 1.4 + 0.6j
    uin, ug = bhs.plane_wave(k=np.asarray(k), direction=np.array([1.0, 0.0, 0.0]))
    kw = dict(uin=uin, uin_grad=ug, k=np.asarray(k), n_end=8, eta=np.asarray(1.0), centers=np.zeros((1, 3)) + 0.25,
              radii=np.array([0.9]), alpha=1.0, beta=0.2j)
    short, full = bhs.biem(c, **kw), bhs.biem(c, force_matrix=True, **kw)
    assert short.matrix is None and rel(short.density, full.density) < 1e-12
    ou, og = bo.plane_wave(k=k, direction=np.array([1.0, 0.0, 0.0]))
    ref = bo.biem("ba", uin=ou, uin_grad=og, k=k, n_end=8, eta=1.0, centers=np.zeros((1, 3)) + 0.25,
                  radii=np.array([0.9]), alpha=1.0, beta=0.2j)
    assert rel(short.density, ref.density) < TOL
    # generic callable == fused plane wave
    gen = bhs.biem(c, uin=lambda x: uin(x), k=np.asarray(k), n_end=6, centers=CEN, radii=RAD)
    fus = bhs.biem(c, uin=uin, k=np.asarray(k), n_end=6, centers=CEN, radii=RAD)
    assert rel(gen.density, fus.density) < 1e-13
    # Im k < 0 warns like the reference (_biem.py:278-285)
    with pytest.warns(UserWarning):
        bhs.biem(c, k=np.asarray(1.0 - 0.1j), n_end=3, centers=CEN, radii=RAD)


@pytest.mark.parametrize("d", [2, 3, 4, 5])
@pytest.mark.parametrize("kind", ["j", "h"])
@pytest.mark.parametrize("derivative", [False, True])
def test_bessel_complex_argument_grid(d, kind, derivative):
    """bhs_bessel_z against scipy (AMOS) over the three algorithm regions of the cylindrical Hankel function
    (J + iY near the real axis, K_nu(-iz) by CF2, Hankel asymptotics) and the spherical closed forms."""
    from biem_helmholtz_sphere_b200 import _ops

    mods = np.array([0.05, 0.3, 1.0, 1.9, 2.1, 3.5, 6.0, 9.0, 14.0, 17.9, 18.1, 25.0, 40.0, 70.0])
    args = np.array([0.0, 0.02, 0.1, 0.3, 0.6, 0.9, 1.2, 1.5, np.pi / 2])
    z = (mods[:, None] * np.exp(1j * args[None, :])).ravel()
    z = z[z.imag != 0]
    n_max = 30
    got = _ops.bessel_z(d, 0 if kind == "j" else 2, n_max, z, derivative).cpu().numpy()  # [nz, n_max+1]
    want = bo.radial(d, n_max, z, kind, derivative=derivative).T
    # compare where the function is representable and scipy itself is meaningful
    ok = np.isfinite(want) & (np.abs(want) > 1e-290) & (np.abs(want) < 1e290)
    err = np.abs(got - want)[ok] / np.abs(want)[ok]
    worst = int(np.argmax(err))
    zi, ni = np.nonzero(ok)[0][worst], np.nonzero(ok)[1][worst]
    print(f"\nbessel_z d={d} {kind} deriv={derivative}: max rel err {err.max():.2e} at z={z[zi]:.3f}, n={ni}")
    # derivatives are differences (n/z) f_n - f_{n+1}: allow the cancellation at zeros of the derivative
    assert err.max() < (1e-10 if not derivative else 1e-8)
    assert np.median(err) < 1e-13


@pytest.mark.parametrize("btype,n_end", [("a", 12), ("bba", 6)])
@pytest.mark.parametrize("k", [1.3 + 0.2j, 0.8 + 1.5j, 3.0 + 4.0j])
def test_complex_k_2d_and_4d(bhs, btype, n_end, k):
    c = bhs.create_from_branching_types(btype)
    d = c.c_ndim
    cen = np.zeros((3, d))
    cen[0, 1], cen[1, 1], cen[2, 0] = 2.0, -2.2, 3.5
    rad = np.array([1.0, 0.8, 1.1])
    dirn = np.zeros(d)
    dirn[0], dirn[1] = 0.6, -0.8
    uin, uin_grad = bhs.plane_wave(k=np.asarray(k), direction=dirn)
    calc = bhs.biem(c, uin=uin, uin_grad=uin_grad, k=np.asarray(k), n_end=n_end, eta=np.asarray(0.9), centers=cen, radii=rad,
                    alpha=1.0, beta=0.2 - 0.1j)
    ou, og = bo.plane_wave(k=k, direction=dirn)
    ref = bo.biem(btype, uin=ou, uin_grad=og, k=k, n_end=n_end, eta=0.9, centers=cen, radii=rad, alpha=1.0, beta=0.2 - 0.1j)
    assert rel(calc.matrix.reshape(ref.matrix.shape), ref.matrix) < TOL
    assert rel(calc.density, ref.density) < TOL
    rng = np.random.default_rng(4)
    x = rng.uniform(-6, 6, size=(d, 40))
    u, want = calc.uscat(x), ref.uscat(x)
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(u), ~ok) and rel(u[ok], want[ok]) < TOL
    xh = x / np.linalg.norm(x, axis=0, keepdims=True)
    assert rel(calc.uscat(xh, far_field=True, per_ball=True), ref.uscat(xh, far_field=True, per_ball=True)) < TOL


def test_point_source_complex_k(bhs):
    c = bhs.create_from_branching_types("ba")
    k = 1.7 + 0.4j
    src = np.array([-5.0, 0.5, 0.2])
    pu, pg = bhs.point_source(k=np.asarray(k), source=src, n=0)
    ou, og = bo.point_source(k=k, source=src, n=0)
    calc = bhs.biem(c, uin=pu, uin_grad=pg, k=np.asarray(k), n_end=9, centers=CEN, radii=RAD, alpha=1.0, beta=0.5)
    ref = bo.biem("ba", uin=ou, uin_grad=og, k=k, n_end=9, centers=CEN, radii=RAD, alpha=1.0, beta=0.5)
    assert rel(calc.density, ref.density) < TOL
