"""Edge cases of the public API on the GPU path: empty point sets, n_end = 1, float32 inputs, CUDA-tensor inputs,
points exactly on a centre / polar axis, uneven radii, and the expand_x switch."""
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


@pytest.mark.parametrize("btype", ["a", "ba", "bba"])
def test_empty_points_and_n_end_one(bhs, btype):
    c = bhs.create_from_branching_types(btype)
    d = c.c_ndim
    cen = bo.grid_centers(0, d)
    k = np.asarray(1.0)
    uin = bhs.plane_wave(k=k, direction=np.eye(d)[0])[0]
    calc = bhs.biem(c, uin=uin, k=k, n_end=1, centers=cen, radii=np.ones(2))
    assert calc.density.shape == (2, 1)
    ref = bo.biem(btype, uin=bo.plane_wave(k=1.0, direction=np.eye(d)[0])[0], k=1.0, n_end=1, centers=cen, radii=np.ones(2))
    assert rel(calc.density, ref.density) < TOL
    assert calc.uscat(np.zeros((d, 0))).shape == (0,)
    assert calc.uscat(np.zeros((d, 0, 4))).shape == (0, 4)
    assert calc.uscat(np.zeros((d, 0)), per_ball=True).shape == (0, 2)
    assert abs(complex(calc.uscat(np.zeros(d))) - complex(ref.uscat(np.zeros(d)))) < TOL


def test_float32_inputs_are_upcast_and_cuda_tensors_stay_on_device(bhs):
    import torch

    c = bhs.create_from_branching_types("ba")
    cen64 = np.array([[0.0, 2.0, 0.0], [0.0, -2.0, 0.0]])
    k = np.float32(1.5)
    uin = bhs.plane_wave(k=np.asarray(k, dtype=np.float32), direction=np.array([1.0, 0.0, 0.0], dtype=np.float32))[0]
    calc = bhs.biem(c, uin=uin, k=np.asarray(k), n_end=6, centers=cen64.astype(np.float32), radii=np.ones(2, np.float32))
    assert calc.density.dtype == np.complex128
    ref = bo.biem("ba", uin=bo.plane_wave(k=1.5, direction=np.array([1.0, 0, 0]))[0], k=1.5, n_end=6, centers=cen64, radii=np.ones(2))
    assert rel(calc.density, ref.density) < TOL
    dev = torch.device("cuda")
    kt = torch.tensor(1.5, dtype=torch.float64, device=dev)
    uin_t = bhs.plane_wave(k=kt, direction=torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device=dev))[0]
    calc_t = bhs.biem(c, uin=uin_t, k=kt, n_end=6, centers=torch.tensor(cen64, device=dev),
                      radii=torch.ones(2, dtype=torch.float64, device=dev))
    assert calc_t.density.is_cuda and calc_t.matrix.is_cuda and calc_t.centers.shape == (3, 2)
    u = calc_t.uscat(torch.zeros(3, 5, dtype=torch.float64, device=dev))
    assert u.is_cuda and u.shape == (5,)
    assert rel(calc_t.density.cpu().numpy(), ref.density) < TOL


def test_points_on_axes_and_centres_uneven_radii(bhs):
    """Degenerate directions: on the polar axis of a ball (sin(theta) = 0) and outside-NaN logic at a ball's centre."""
    c = bhs.create_from_branching_types("ba")
    cen = np.array([[0.0, 2.0, 0.0], [0.0, -2.5, 0.0], [5.0, 0.0, 0.0]])
    rad = np.array([1.0, 1.4, 0.3])
    k = np.asarray(2.0)
    calc = bhs.biem(c, uin=bhs.plane_wave(k=k, direction=np.array([0.0, 1.0, 0.0]))[0], k=k, n_end=12, centers=cen, radii=rad)
    ref = bo.biem("ba", uin=bo.plane_wave(k=2.0, direction=np.array([0.0, 1.0, 0.0]))[0], k=2.0, n_end=12, centers=cen, radii=rad)
    x = np.array([[3.0, 0.0, 0.0],     # on the polar (x0) axis of ball 2, between balls
                  [-4.0, 2.0, 0.0],    # on the polar axis through ball 0's centre
                  [0.0, 2.0, 0.0],     # centre of ball 0 -> NaN (inside)
                  [0.0, -2.5, 1.4 + 1e-12],  # just outside ball 1
                  [5.0, 0.0, 0.3 - 1e-12]]).T  # just inside ball 2 -> NaN
    u = calc.uscat(x)
    want = ref.uscat(x)
    assert np.array_equal(np.isnan(u), np.isnan(want))
    assert np.array_equal(np.isnan(u), [False, False, True, False, True])
    ok = ~np.isnan(want)
    assert rel(u[ok], want[ok]) < TOL


def test_expand_x_false_matches_loop(bhs):
    c = bhs.create_from_branching_types("a")
    cen = np.array([[0.0, 2.0], [0.0, -2.0]])
    ks = np.array([0.8, 1.9, 3.1])
    uin = bhs.plane_wave(k=ks, direction=np.array([[1.0], [0.0]]))[0]
    calc = bhs.biem(c, uin=uin, k=ks, n_end=9, centers=cen[None], radii=np.ones((1, 2)), eta=np.ones(3))
    rng = np.random.default_rng(5)
    x = rng.uniform(3.5, 6.0, size=(2, 7, 3))  # one point set per system
    u = calc.uscat(x, expand_x=False)
    assert u.shape == (7, 3)
    ue = calc.uscat(x[:, :, 1])  # expand: same points for every system
    assert ue.shape == (7, 3)
    assert np.allclose(u[:, 1], ue[:, 1], rtol=1e-13, atol=0)
    for i, k in enumerate(ks):
        ref = bo.biem("a", uin=bo.plane_wave(k=k, direction=np.array([1.0, 0.0]))[0], k=k, n_end=9, centers=cen, radii=np.ones(2))
        assert rel(u[:, i], ref.uscat(x[:, :, i])) < TOL


def test_batched_geometries_and_per_ball_alpha(bhs):
    """Leading batch axis on centers / radii (a different geometry per system): identical to a loop of calls.  Per-ball
    alpha / beta arrays go with a scalar k -- the reference's own shape check (_biem.py:292-306 broadcasts alpha.shape,
    ball axis included, against k.shape) rejects them together with a batched k, and so does the mirror."""
    c = bhs.create_from_branching_types("ba")
    rng = np.random.default_rng(11)
    K, B, n_end = 3, 2, 7
    cen = np.stack([np.array([[0.0, 2.0 + 0.3 * i, 0.0], [0.4 * i, -2.0, 0.2]]) for i in range(K)])
    rad = np.stack([np.array([1.0 - 0.1 * i, 0.8 + 0.1 * i]) for i in range(K)])
    ks = np.array([0.9, 1.7, 2.6])
    uin, uin_grad = bhs.plane_wave(k=ks, direction=np.array([[1.0], [0.0], [0.0]]))
    calc = bhs.biem(c, uin=uin, uin_grad=uin_grad, k=ks, n_end=n_end, eta=np.ones(K), centers=cen, radii=rad, alpha=0.7,
                    beta=0.2 + 0.1j)
    assert calc.density.shape == (K, B, n_end * n_end) and calc.matrix.shape == (K, B, n_end * n_end, B, n_end * n_end)
    x = np.array([[0.0, 5.0], [0.0, 0.5], [0.0, -1.0]])
    u = calc.uscat(x)
    for i in range(K):
        ou, og = bo.plane_wave(k=ks[i], direction=np.array([1.0, 0.0, 0.0]))
        ref = bo.biem("ba", uin=ou, uin_grad=og, k=ks[i], n_end=n_end, eta=1.0, centers=cen[i], radii=rad[i], alpha=0.7,
                      beta=0.2 + 0.1j)
        assert rel(calc.density[i], ref.density) < TOL
        assert rel(calc.matrix[i], ref.matrix) < TOL
        assert rel(u[:, i], ref.uscat(x)) < TOL
    alpha = rng.normal(size=B) + 1j * rng.normal(size=B)
    beta = rng.normal(size=B) + 1j * rng.normal(size=B)
    k = np.asarray(1.3)
    uin, uin_grad = bhs.plane_wave(k=k, direction=np.array([1.0, 0.0, 0.0]))
    one = bhs.biem(c, uin=uin, uin_grad=uin_grad, k=k, n_end=n_end, centers=cen[1], radii=rad[1], alpha=alpha, beta=beta)
    ou, og = bo.plane_wave(k=1.3, direction=np.array([1.0, 0.0, 0.0]))
    ref = bo.biem("ba", uin=ou, uin_grad=og, k=1.3, n_end=n_end, centers=cen[1], radii=rad[1], alpha=alpha, beta=beta)
    assert rel(one.density, ref.density) < TOL
    with pytest.raises(ValueError):
        bhs.biem(c, uin=uin, k=ks, n_end=n_end, eta=np.ones(K), centers=cen, radii=rad, alpha=np.ones((K, B)))


def test_hand_built_result_record(bhs):
    """A BIEMResultCalculator assembled by hand (e.g. from a density computed elsewhere) evaluates like the oracle."""
    c = bhs.create_from_branching_types("ba")
    cen = np.array([[0.0, 2.0, 0.0], [0.0, -2.0, 0.0], [3.5, 0.0, 1.0]])
    rad = np.array([1.0, 0.9, 0.6])
    ref = bo.biem("ba", uin=bo.plane_wave(k=1.9, direction=np.array([0.0, 0.0, 1.0]))[0], k=1.9, n_end=8, eta=0.6,
                  centers=cen, radii=rad)
    rec = bhs.BIEMResultCalculator(c=c, centers=cen.T.copy(), radii=rad, k=np.asarray(1.9), n_end=8, eta=np.asarray(0.6),
                                   kind="outer", density=ref.density)
    rng = np.random.default_rng(2)
    x = rng.uniform(-5, 5, size=(3, 4, 6))
    u, want = rec.uscat(x), ref.uscat(x)
    ok = ~np.isnan(want)
    assert u.shape == (4, 6) and np.array_equal(np.isnan(u), ~ok)
    assert rel(u[ok], want[ok]) < TOL
    assert rel(rec.uscat(x, per_ball=True)[ok], ref.uscat(x, per_ball=True)[ok]) < TOL


@pytest.mark.parametrize("k", [2.3, 1.1 + 0.7j])
@pytest.mark.parametrize("case", ["coplanar", "points_off_plane", "centres_off_plane", "negative_side"])
def test_uscat_planar_fast_path_selection(bhs, case, k):
    """The 3-D field kernel has a device-selected fast path for points and centres that all share one x2 (planar heat
    maps).  Both selections must agree with the oracle, including points on both sides of a ball (azimuth 0 and pi)."""
    from biem_helmholtz_sphere_b200 import _ops

    rng = np.random.default_rng(8)
    B, n_end = 3, 13
    cen = np.array([[0.0, 2.0, 0.4], [0.5, -2.0, 0.4], [4.0, 0.3, 0.4]])
    if case == "centres_off_plane":
        cen[1, 2] = 0.1
    rad = np.array([1.0, 0.8, 1.2])
    deg = bo.degree_table("ba", n_end)
    dens = (rng.normal(size=(B, n_end * n_end)) + 1j * rng.normal(size=(B, n_end * n_end))) * (0.6 ** deg)[None, :]
    x = rng.uniform(-6, 6, size=(3, 257))
    x[2] = 0.4
    if case == "points_off_plane":
        x[2, 100] = 0.4000001
    if case == "negative_side":
        x[1] = -np.abs(x[1]) - 3.5  # every point has dx1 < 0 for every ball
    x[:, 0] = [0.0, 2.0 + 2.0, 0.4]  # on the in-plane axis through ball 0's centre
    res = bo.OracleResult(c=bo.OracleCoordinates("ba"), centers=cen.T.copy(), radii=rad, k=k, n_end=n_end, eta=0.8,
                          kind="outer", density=dens, matrix=None)
    for pb in (False, True):
        want = bo.biem_u(res, x, per_ball=pb)
        got = _ops.uscat(3, n_end, cen, rad, k, 0.8, dens, x, per_ball=pb).cpu().numpy()
        nan = np.isnan(want)
        assert np.array_equal(nan, np.isnan(got))
        assert rel(got[~nan], want[~nan]) < 1e-11


@pytest.mark.parametrize("n_end", [1, 3, 4, 8, 13, 17, 20, 24, 29, 32])
@pytest.mark.parametrize("k", [1.3, 0.9 + 0.4j])
def test_uscat_planar_kernel_every_band(bhs, n_end, k):
    """The planar field kernel (coefficients rotated into the frame whose polar axis is the plane's normal) is instantiated
    per band of four degrees; every band, real and complex k, outer and inner masks, against the oracle at 1e-11.  The plane
    is x2 = -0.7 (not a coordinate plane of the centres' frame origin) and the points surround the balls (all azimuths)."""
    from biem_helmholtz_sphere_b200 import _ops

    rng = np.random.default_rng(80 + n_end)
    B = 4
    cen = np.array([[0.0, 2.0, -0.7], [0.5, -2.0, -0.7], [4.0, 0.3, -0.7], [-3.5, -0.4, -0.7]])
    rad = np.array([1.0, 0.8, 1.2, 0.9])
    deg = bo.degree_table("ba", n_end)
    dens = (rng.normal(size=(B, n_end * n_end)) + 1j * rng.normal(size=(B, n_end * n_end))) * (0.55 ** deg)[None, :]
    x = rng.uniform(-7, 7, size=(3, 300))
    x[2] = -0.7
    x[:, 0] = [0.0 + 2.5, 2.0, -0.7]   # on the x0 line through ball 0's centre (phi' = 0)
    x[:, 1] = [0.0, 2.0 - 2.5, -0.7]   # phi' = -pi/2
    res = bo.OracleResult(c=bo.OracleCoordinates("ba"), centers=cen.T.copy(), radii=rad, k=k, n_end=n_end, eta=0.8,
                          kind="outer", density=dens, matrix=None)
    for pb in (False, True):
        want = bo.biem_u(res, x, per_ball=pb)
        got = _ops.uscat(3, n_end, cen, rad, k, 0.8, dens, x, per_ball=pb).cpu().numpy()
        nan = np.isnan(want)
        assert np.array_equal(nan, np.isnan(got)) and nan.any() and not nan.all()
        assert rel(got[~nan], want[~nan]) < 1e-11
    # inner problem: one ball, points inside it, in its equatorial plane
    xi = cen[0][:, None] + np.vstack([rng.uniform(-0.6, 0.6, size=(2, 50)), np.zeros((1, 50))])
    res_i = bo.OracleResult(c=bo.OracleCoordinates("ba"), centers=cen[:1].T.copy(), radii=rad[:1], k=k, n_end=n_end, eta=0.8,
                            kind="inner", density=dens[:1], matrix=None)
    want = bo.biem_u(res_i, xi)
    got = _ops.uscat(3, n_end, cen[:1], rad[:1], k, 0.8, dens[:1], xi, inner=True).cpu().numpy()
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(got), ~ok)
    assert float(np.max(np.abs(got[ok] - want[ok]) / np.abs(want[ok]))) < 1e-10


@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (0.0, 1.0), (1.0, 0.5 + 0.2j)])
def test_boundary_condition_residual_on_the_spheres(bhs, alpha, beta):
    """Physics check that does not go through the oracle: the total field u_in + u_s of the GPU solution must
    satisfy alpha u + beta du/dn = 0 on every sphere.  The normal derivative is a one-sided 4-point finite difference of
    the GPU field, so the residual is limited by the difference formula (h^3), not by the solver."""
    c = bhs.create_from_branching_types("ba")
    cen = np.array([[0.0, 2.0, 0.0], [0.3, -2.1, 0.4], [3.6, 0.0, -0.5]])
    rad = np.array([1.0, 0.8, 1.1])
    k = np.asarray(1.3)
    dirn = np.array([0.6, 0.0, 0.8])
    uin, uin_grad = bhs.plane_wave(k=k, direction=dirn)
    calc = bhs.biem(c, uin=uin, uin_grad=uin_grad, k=k, n_end=22, centers=cen, radii=rad, alpha=alpha, beta=beta,
                    keep_matrix=False)
    rng = np.random.default_rng(9)
    y = rng.normal(size=(3, 40))
    y /= np.linalg.norm(y, axis=0, keepdims=True)
    h = 2e-3
    worst = 0.0
    for b in range(3):
        f = []
        for j in range(4):
            x = cen[b][:, None] + rad[b] * (1.0 + 1e-9 + j * h) * y  # 1e-9: stay outside the NaN mask (r < rho) despite rounding
            f.append(calc.uscat(x) + uin(x))
        dn = (-11.0 * f[0] + 18.0 * f[1] - 9.0 * f[2] + 2.0 * f[3]) / (6.0 * h * rad[b])
        res = alpha * f[0] + beta * dn
        assert not np.any(np.isnan(res))
        worst = max(worst, float(np.max(np.abs(res))))
    print(f"\nboundary condition residual alpha={alpha} beta={beta}: {worst:.2e}")
    # Dirichlet: floor = aliasing error of the reference's n_end-point RHS quadrature (4.7e-9 here); with beta != 0 the
    # one-sided difference formula dominates (1.8e-7 / 1.7e-6 measured)
    assert worst < (5e-8 if beta == 0.0 else 1e-5)


@pytest.mark.parametrize("btype,relabel", [("bpa", True), ("bpa", False), ("bpbpa", True), ("caa", False), ("bbpa", True)])
def test_other_trees_generic_callables_and_frames(bhs, btype, relabel):
    """b' / c trees through the public API against the oracle, with a NON-tagged incident field (the boundary
    data is sampled by calling back into Python: the quadrature points and normals must reach the callables in the caller's
    cartesian frame, not in the chain frame the device code works in), Robin data, off-axis geometry and a field evaluation
    away from the origin -- none of which the golden rows (symmetric geometry, origin only) can see."""
    c = bhs.create_from_branching_types(btype)
    o = bo.OracleCoordinates(btype)
    d = c.c_ndim
    if relabel:
        c, o = c.relabel({0: d - 1, d - 1: 0}), o.relabel({0: d - 1, d - 1: 0})
    rng = np.random.default_rng(len(btype) + d)
    cen = rng.normal(size=(2, d)) * 0.4
    cen[0, 0] += 2.2
    cen[1, 1] -= 2.3
    rad = np.array([1.0, 0.8])
    dirn = rng.normal(size=d)
    dirn /= np.linalg.norm(dirn)
    k, n_end = 1.2, 5
    alpha, beta = 1.0, 0.3 + 0.1j

    def uin(x):
        dd = dirn[(slice(None),) + (None,) * (x.ndim - 1)]
        return np.exp(1j * k * np.sum(dd * x, axis=0))

    def uin_grad(x):
        dd = dirn[(slice(None),) + (None,) * (x.ndim - 1)]
        return 1j * k * dd * uin(x)[None]

    calc = bhs.biem(c, uin=uin, uin_grad=uin_grad, k=np.asarray(k), n_end=n_end, eta=np.asarray(0.9), centers=cen, radii=rad,
                    alpha=alpha, beta=beta, keep_matrix=False)
    ref = bo.biem(o, centers=cen, radii=rad, k=k, n_end=n_end, uin=uin, uin_grad=uin_grad, eta=0.9, alpha=alpha, beta=beta)
    x = rng.uniform(-4, 4, size=(d, 60))
    u, ur = np.asarray(calc.uscat(x)), ref.uscat(x)
    ok = ~np.isnan(ur)
    assert np.array_equal(np.isnan(u), ~ok) and ok.sum() > 20
    assert rel(u[ok], ur[ok]) < TOL
    # the tagged plane wave (fused right-hand side kernel) must agree with the callback path
    pw, pwg = bhs.plane_wave(k=np.asarray(k), direction=dirn)
    calc2 = bhs.biem(c, uin=pw, uin_grad=pwg, k=np.asarray(k), n_end=n_end, eta=np.asarray(0.9), centers=cen, radii=rad,
                     alpha=alpha, beta=beta, keep_matrix=False)
    assert rel(np.asarray(calc2.uscat(x))[ok], ur[ok]) < TOL
    assert np.array_equal(np.asarray(calc.centers), cen.T)  # stored as given (transposed), in the caller's frame


def test_evolved_record_uses_its_public_fields(bhs):
    """attrs.evolve(res, density=...) must evaluate the NEW density (the device-side cache of biem() is not carried over),
    as the reference's biem_u, which always reads the public fields, does."""
    import attrs

    c = bhs.create_from_branching_types("ba")
    cen = bo.grid_centers(0, 3)
    k = np.asarray(1.0)
    uin = bhs.plane_wave(k=k, direction=np.eye(3)[0])[0]
    calc = bhs.biem(c, uin=uin, k=k, n_end=5, centers=cen, radii=np.ones(2), keep_matrix=False)
    x = np.array([[0.3, 3.0], [0.1, -0.5], [2.5, 0.7]])
    u1 = np.asarray(calc.uscat(x))
    doubled = attrs.evolve(calc, density=2.0 * np.asarray(calc.density))
    assert rel(np.asarray(doubled.uscat(x)), 2.0 * u1) < 1e-13
    moved = attrs.evolve(calc, radii=np.array([0.5, 0.5]))
    ref = bo.OracleResult(c=bo.OracleCoordinates("ba"), centers=cen.T.copy(), radii=np.array([0.5, 0.5]), k=1.0, n_end=5,
                          eta=1.0, kind="outer", density=np.asarray(calc.density), matrix=None)
    assert rel(np.asarray(moved.uscat(x)), ref.uscat(x)) < TOL


def test_singular_system_raises_or_poisons(bhs):
    """An exactly singular system (alpha = beta = 0: the whole matrix vanishes): LinAlgError for NumPy callers, as the
    reference's solve raises; NaN densities (no host synchronisation) for torch callers."""
    import torch

    c = bhs.create_from_branching_types("ba")
    cen = bo.grid_centers(0, 3)
    k = np.asarray(1.0)
    uin = bhs.plane_wave(k=k, direction=np.eye(3)[0])[0]
    with pytest.raises(np.linalg.LinAlgError):
        bhs.biem(c, uin=uin, k=k, n_end=4, centers=cen, radii=np.ones(2), alpha=0.0, beta=0.0, keep_matrix=False)
    kt = torch.tensor([1.0, 2.0], dtype=torch.float64, device="cuda")
    uin_t = bhs.plane_wave(k=kt, direction=torch.tensor([[1.0], [0.0], [0.0]], dtype=torch.float64, device="cuda"))[0]
    res = bhs.biem(c, uin=uin_t, k=kt, n_end=4, centers=torch.as_tensor(cen, device="cuda")[None],
                   radii=torch.ones(1, 2, dtype=torch.float64, device="cuda"), alpha=0.0, beta=0.0, keep_matrix=False)
    assert bool(torch.isnan(res.density.real).all())


def test_pinned_host_tensors_in_and_out(bhs):
    """Pinned torch CPU tensors in -> pinned torch CPU tensors out (the tiled host pipeline of large field evaluations and
    the asynchronous result copies), same values as the NumPy path."""
    import torch

    c = bhs.create_from_branching_types("ba")
    cen = bo.grid_centers(1, 3)
    k = np.asarray(1.1)
    uin = bhs.plane_wave(k=k, direction=np.eye(3)[0])[0]
    calc = bhs.biem(c, uin=uin, k=k, n_end=8, centers=cen, radii=np.ones(4), keep_matrix=False)
    g = np.linspace(-9, 9, 800)
    x = np.zeros((3, 800, 800))
    x[0], x[1] = g[:, None], g[None, :]
    x[2] = 0.25  # off the plane of the centres: the general kernel, tiled over the three copy / compute streams
    xp = torch.as_tensor(x).pin_memory()
    u_pin = calc.uscat(xp)
    assert isinstance(u_pin, torch.Tensor) and u_pin.device.type == "cpu" and u_pin.is_pinned() and u_pin.shape == (800, 800)
    u_np = np.asarray(calc.uscat(x))
    assert np.array_equal(np.isnan(u_np), np.isnan(u_pin.numpy()))
    ok = ~np.isnan(u_np)
    assert np.array_equal(u_np[ok], u_pin.numpy()[ok])
    sub = (slice(None), slice(0, 800, 37), slice(0, 800, 41))
    ref = bo.OracleResult(c=bo.OracleCoordinates("ba"), centers=cen.T.copy(), radii=np.ones(4), k=1.1, n_end=8, eta=1.0,
                          kind="outer", density=np.asarray(calc.density), matrix=None)
    want = ref.uscat(x[sub])
    okw = ~np.isnan(want)
    assert rel(u_np[sub[1:]][okw], want[okw]) < TOL


def test_complex_wavenumber_with_nonpositive_real_part(bhs):
    """Im k > 0 makes any Re k a valid argument of h^(1)_n (absorbing / evanescent media): bhs_uscat accepts it (3-D)."""
    from biem_helmholtz_sphere_b200 import _ops

    rng = np.random.default_rng(3)
    n_end, B = 9, 2
    cen = np.array([[0.0, 2.0, 0.1], [0.3, -2.0, -0.2]])
    rad = np.array([1.0, 0.9])
    dens = (rng.normal(size=(B, n_end * n_end)) + 1j * rng.normal(size=(B, n_end * n_end))) * (0.5 ** bo.degree_table("ba", n_end))
    x = rng.uniform(-5, 5, size=(3, 200))
    for k in (-0.4 + 1.1j, 0.0 + 0.8j):
        res = bo.OracleResult(c=bo.OracleCoordinates("ba"), centers=cen.T.copy(), radii=rad, k=k, n_end=n_end, eta=0.7,
                              kind="outer", density=dens, matrix=None)
        want = bo.biem_u(res, x)
        got = _ops.uscat(3, n_end, cen, rad, k, 0.7, dens, x).cpu().numpy()
        ok = ~np.isnan(want)
        assert np.array_equal(np.isnan(got), ~ok) and rel(got[ok], want[ok]) < TOL
    with pytest.raises(NotImplementedError):
        _ops.uscat(3, n_end, cen, rad, -1.0, 0.7, dens, x)
