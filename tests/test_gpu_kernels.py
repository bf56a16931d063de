"""GPU parity tests of the individual kernels, called through the C ABI (ctypes), against the CPU oracle.

Tolerances are written per test; floating point throughout (complex128), index tables bit-exact.
"""
import numpy as np
import pytest

from oracle import biem_oracle as bo

pytestmark = pytest.mark.gpu

BT = {2: "a", 3: "ba", 4: "bba"}


@pytest.fixture(scope="module")
def ops():
    import torch

    assert torch.cuda.is_available()
    from biem_helmholtz_sphere_b200 import _ops

    return _ops


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_fp64_peaks(ops):
    dfma = ops.fp64_peak(0, 2048)
    dmma = ops.fp64_peak(1, 2048)
    print(f"\nFP64 peaks: DFMA {dfma:.2f} TFLOP/s, DMMA(m8n8k4) {dmma:.2f} TFLOP/s")
    assert dfma > 1.0 and dmma > 1.0


@pytest.mark.parametrize("d", [2, 3, 4, 5])
@pytest.mark.parametrize("kind", ["j", "y", "h"])
@pytest.mark.parametrize("derivative", [False, True])
def test_bessel(ops, d, kind, derivative):
    n_max = 40
    x = np.concatenate([np.geomspace(1e-2, 300.0, 400), np.linspace(0.3, 60.0, 333), [0.5, 1.0, 8.0, 24.9, 25.0, 25.1]])
    got = ops.bessel(d, {"j": 0, "y": 1, "h": 2}[kind], n_max, x, derivative).cpu().numpy()
    want = bo.radial(d, n_max, x, kind, derivative).T  # [nx, n]
    with np.errstate(all="ignore"):
        err = np.abs(got - want) / np.abs(want)
    finite = np.isfinite(want) & (np.abs(want) > 1e-290) & (np.abs(want) < 1e290)
    # near zeros of an oscillating function only absolute accuracy (relative to the local envelope |h|) is
    # meaningful; in the monotone region (order > argument) the error is purely relative
    if kind != "h":
        env = np.abs(bo.radial(d, n_max, x, "h", derivative).T)
        osc = x[:, None] > (np.arange(n_max + 1)[None, :] + d / 2.0 - 1.0)
        denom = np.where(osc & np.isfinite(env), np.maximum(np.abs(want), 0.1 * env), np.abs(want))
        with np.errstate(all="ignore"):
            err = np.abs(got - want) / denom
    worst = float(np.nanmax(np.where(finite, err, 0.0)))
    print(f"\nbessel d={d} kind={kind} deriv={derivative}: max rel err {worst:.2e}")
    assert worst < 2e-12


@pytest.mark.parametrize("d,n_end", [(2, 9), (2, 70), (3, 6), (3, 24), (4, 5), (4, 10)])
def test_harmonics(ops, d, n_end):
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(d, 257))
    xyz[:, 0] = 0.0
    xyz[0, 0] = 1.0  # on the polar axis
    xyz[:, 1] = 0.0
    xyz[d - 1, 1] = -2.0
    c = bo.OracleCoordinates(BT[d])
    sph = c.from_cartesian(xyz)
    want = bo.harmonics(BT[d], [sph[i] for i in range(d - 1)], n_end)
    got = ops.harmonics(d, n_end, xyz).cpu().numpy()
    err = np.max(np.abs(got - want))
    print(f"\nharmonics d={d} n_end={n_end}: max abs err {err:.2e}")
    assert err < 5e-13
    got2 = ops.harmonics(d, n_end, xyz, double_band=True).cpu().numpy()
    want2 = bo.harmonics(BT[d], [sph[i] for i in range(d - 1)], 2 * n_end - 1)
    assert np.max(np.abs(got2 - want2)) < 2e-12


@pytest.mark.parametrize("d,n_end", [(2, 7), (3, 6), (3, 16), (4, 6)])
def test_plan_tables_bit_exact(ops, d, n_end):
    from biem_helmholtz_sphere_b200._lib import get_plan

    plan = get_plan(d, n_end)
    tab = bo.index_tables(BT[d], n_end)
    assert plan.H == tab.shape[0] == bo.harm_count(d, n_end)
    assert np.array_equal(plan.index_table().astype(np.int64), tab)
    dirs, w = plan.quadrature()
    angles, w0 = bo.quadrature(BT[d], n_end)
    c = bo.OracleCoordinates(BT[d])
    dirs0 = c.to_cartesian({i: a for i, a in enumerate(angles)})
    # same node set (the b-node order may be reversed: the rule is symmetric) -> compare as sorted sets
    key = np.lexsort(np.round(dirs, 9))
    key0 = np.lexsort(np.round(dirs0, 9))
    assert np.max(np.abs(dirs[:, key] - dirs0[:, key0])) < 1e-13
    assert np.max(np.abs(w[key] - w0[key0])) < 1e-14


@pytest.mark.parametrize("d,n_end,B", [(2, 8, 3), (3, 6, 2), (3, 12, 4), (4, 5, 2)])
def test_rhs_expand(ops, d, n_end, B):
    rng = np.random.default_rng(1)
    centers = rng.normal(size=(B, d)) * 3
    radii = rng.uniform(0.5, 1.5, size=B)
    direction = np.zeros(d)
    direction[0] = 0.6
    direction[1] = 0.8
    alpha = rng.normal(size=B) + 1j * rng.normal(size=B)
    beta = rng.normal(size=B) + 1j * rng.normal(size=B)
    ks = np.array([0.7, 1.9])
    want = []
    for k in ks:
        uin, uin_grad = bo.plane_wave(k=k, direction=direction)
        g = bo.boundary_data(BT[d], centers, radii, n_end, alpha, beta, uin, uin_grad)
        want.append(bo.expand(BT[d], g, n_end))
    want = np.stack(want)
    got = ops.rhs_expand(d, n_end, centers=centers, radii=radii, k_in=ks, direction=direction, alpha=alpha, beta=beta)
    e1 = relerr(got.cpu().numpy(), want)
    # sampled-data path: feed the oracle's own g through the plan's node order
    from biem_helmholtz_sphere_b200._lib import get_plan

    dirs, _ = get_plan(d, n_end).quadrature()
    gs = []
    for k in ks:
        uin, uin_grad = bo.plane_wave(k=k, direction=direction)
        x = radii[None, None, :] * dirs[:, :, None] + centers.T[:, None, :]
        g = -alpha[None, :] * uin(x) - beta[None, :] * np.sum(uin_grad(x) * dirs[:, :, None], axis=0)
        gs.append(g)
    got2 = ops.rhs_expand(d, n_end, g=np.stack(gs))
    e2 = relerr(got2.cpu().numpy(), want)
    print(f"\nrhs d={d} n_end={n_end}: fused {e1:.2e} sampled {e2:.2e}")
    assert e1 < 1e-12 and e2 < 1e-12


@pytest.mark.parametrize("d,n_end,B", [(2, 6, 3), (2, 32, 4), (3, 4, 2), (3, 6, 3), (3, 10, 4), (3, 16, 2), (4, 3, 2), (4, 6, 3)])
def test_assemble(ops, d, n_end, B):
    rng = np.random.default_rng(2)
    centers = np.zeros((B, d))
    centers[:, 0] = 3.1 * np.arange(B)
    centers[:, 1:] = rng.normal(size=(B, d - 1)) * 0.7
    radii = rng.uniform(0.6, 1.2, size=B)
    alpha = rng.normal(size=B) + 1j * rng.normal(size=B)
    beta = rng.normal(size=B) + 1j * rng.normal(size=B)
    ks = np.array([0.8, 2.3])
    etas = np.array([1.0, 0.4])
    got = ops.assemble(d, n_end, centers, radii, ks, etas, alpha, beta).cpu().numpy()
    for i, (k, eta) in enumerate(zip(ks, etas)):
        want = bo.assemble(BT[d], centers, radii, float(k), n_end, float(eta), alpha, beta)
        N = want.shape[0] * want.shape[1]
        want = want.reshape(N, N)
        # entries span many orders of magnitude: compare in the row/column scaling of the diagonal
        sc = np.sqrt(np.abs(np.diag(want)))
        e = np.max(np.abs(got[i] - want) / (sc[:, None] * sc[None, :]))
        e_abs = relerr(got[i], want)
        print(f"\nassemble d={d} n_end={n_end} B={B} k={k}: scaled err {e:.2e}, max-norm rel {e_abs:.2e}")
        assert e_abs < 1e-12


@pytest.mark.parametrize("d,n_end,B", [(2, 9, 5), (3, 7, 4), (4, 4, 3)])
def test_assemble_rows_strips(ops, d, n_end, B):
    """bhs_assemble_rows: the strips of any split of the row balls are bit-identical to the rows of the full matrix."""
    import torch

    rng = np.random.default_rng(d)
    centers = np.zeros((B, d))
    centers[:, 1] = 4.0 * np.arange(B) - 2.0 * (B - 1)
    centers[:, 0] = rng.uniform(-0.3, 0.3, size=B)
    radii = rng.uniform(0.7, 1.1, size=B)
    k, eta = np.array([1.3]), np.array([0.8])
    al = rng.normal(size=B) + 1j * rng.normal(size=B)
    be = rng.normal(size=B) + 1j * rng.normal(size=B)
    full = ops.assemble(d, n_end, centers, radii, k, eta, al, be)[0]
    H = full.shape[0] // B
    for b_lo, b_hi in ((0, 1), (1, B), (B - 1, B), (0, B), (1, 2)):
        strip = ops.assemble(d, n_end, centers, radii, k, eta, al, be, rows=(b_lo, b_hi))[0]
        assert strip.shape == ((b_hi - b_lo) * H, B * H)
        assert torch.equal(strip, full[b_lo * H : b_hi * H])


@pytest.mark.parametrize("d,half,n_end,nsys", [(3, 2, 16, 2), (3, 2, 16, 4), (3, 1, 7, 3), (2, 2, 20, 2), (3, 3, 10, 2), (3, 4, 6, 1)])
def test_assemble_kernels_agree_bitwise(ops, monkeypatch, d, half, n_end, nsys):
    """The register-resident assembly kernel against the shared-memory-resident one it replaced (BHS_ASM_LEGACY=1), on
    regular grids (shared translations, many CTA iterations): bit-identical without the merging of opposite translations
    (BHS_ASM_NOPM=1) and identical as numbers with it; repeated, because a race between warps of a CTA would show up as an
    occasional difference (one did: the member list of the next translation overwrote the one still being read)."""
    import torch

    from biem_helmholtz_sphere_b200.geometry import grid_centers

    cen = torch.as_tensor(grid_centers(half, d), device="cuda")
    B = cen.shape[0]
    g = torch.Generator(device="cuda").manual_seed(3)
    rad = 0.5 + 0.4 * torch.rand(B, dtype=torch.float64, device="cuda", generator=g)
    k = torch.linspace(0.7, 2.9, nsys, dtype=torch.float64, device="cuda")
    monkeypatch.setenv("BHS_ASM_LEGACY", "1")
    want = ops.assemble(d, n_end, cen, rad, k, k)
    monkeypatch.delenv("BHS_ASM_LEGACY")
    for rep in range(4):
        monkeypatch.setenv("BHS_ASM_NOPM", "1")
        got = ops.assemble(d, n_end, cen, rad, k, k)
        assert torch.equal(got.view(torch.float64).view(torch.int64), want.view(torch.float64).view(torch.int64)), rep
        monkeypatch.delenv("BHS_ASM_NOPM")
        got = ops.assemble(d, n_end, cen, rad, k, k)
        assert torch.equal(got, want), rep  # equal as numbers (signed zeros may differ)


@pytest.mark.parametrize("M,N,K", [(64, 64, 8), (128, 192, 32), (200, 77, 40), (1000, 520, 128), (37, 5, 16)])
def test_zgemm_sub(ops, M, N, K):
    import torch

    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(M, K, dtype=torch.complex128, device="cuda", generator=g)
    B = torch.randn(K, N, dtype=torch.complex128, device="cuda", generator=g)
    Cm = torch.randn(M, N, dtype=torch.complex128, device="cuda", generator=g)
    want = Cm - A @ B
    got = ops.zgemm_sub_(Cm.clone(), A, B)
    err = float((got - want).abs().max() / want.abs().max())
    print(f"\nzgemm {M}x{N}x{K}: rel err {err:.2e}")
    assert err < 1e-13


@pytest.mark.parametrize("N", [8, 31, 32, 33, 72, 126, 128, 129, 200, 770, 1000, 2048])
def test_zgesv_random(ops, N):
    import torch

    g = torch.Generator(device="cuda").manual_seed(N)
    A = torch.randn(N, N, dtype=torch.complex128, device="cuda", generator=g)
    b = torch.randn(N, dtype=torch.complex128, device="cuda", generator=g)
    want = torch.linalg.solve(A, b)
    x, bufs = ops.zgesv_(A.clone(), b.clone())
    torch.cuda.synchronize()
    assert int(bufs.info.item()) == 0
    res = float((A @ x - b).abs().max() / (A.abs().max() * x.abs().max() * N))
    err = float((x - want).abs().max() / want.abs().max())
    print(f"\nzgesv N={N}: err vs torch {err:.2e}, scaled residual {res:.2e}")
    assert res < 1e-14
    assert err < 1e-8  # random matrices are not well conditioned; the residual is the real check


@pytest.mark.parametrize("N,S,nrhs", [(33, 3, 1), (200, 5, 2), (515, 4, 1), (1024, 8, 1)])
def test_zgesv_batched_equals_single(ops, N, S, nrhs):
    """bhs_zgesv_batched (S systems in lock step, system index in blockIdx.z) against S separate bhs_zgesv calls; systems
    get different pivot orders (rows of system s are scaled / permuted differently).  A lone system picks its pivots inside
    one thread-block cluster (plain partial pivoting), systems in lock step by tournament: both are valid LU factorisations
    of the same matrix, so the solutions agree to rounding (and are bit-identical when the tournament serves both)."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(100 + N)
    A = torch.randn(S, N, N, dtype=torch.complex128, device="cuda", generator=g)
    for s_ in range(S):
        A[s_] = A[s_][torch.randperm(N, device="cuda", generator=g)] * (10.0 ** (s_ - 2))
    b = torch.randn((S, N) if nrhs == 1 else (S, N, nrhs), dtype=torch.complex128, device="cuda", generator=g)
    xb, bufs = ops.zgesv_batched_(A.clone(), b.clone())
    torch.cuda.synchronize()
    assert bool(torch.all(bufs.info == 0))
    for s_ in range(S):
        x1, b1 = ops.zgesv_(A[s_].clone(), b[s_].clone())
        dev_ = float((xb[s_] - x1).abs().max() / x1.abs().max())
        assert dev_ < 1e-11, f"system {s_} differs from the single-system solve by {dev_:.2e}"
        # both pivot sequences are permutations of 0..N-1 when replayed as row swaps
        for piv in (bufs.ipiv[s_ * N:(s_ + 1) * N], b1.ipiv):
            assert int(piv.min()) >= 0 and int(piv.max()) < N and bool(torch.all(piv >= torch.arange(N, device="cuda")))
        rhs_s = b[s_] if nrhs > 1 else b[s_][:, None]
        xs = xb[s_] if nrhs > 1 else xb[s_][:, None]
        res = float((A[s_] @ xs - rhs_s).abs().max() / (A[s_].abs().max() * xs.abs().max() * N))
        assert res < 1e-14


@pytest.mark.parametrize("N,S", [(1000, 2), (2304, 3)])
def test_zgesv_batched_block_widths(ops, monkeypatch, N, S):
    """Systems factorised in lock step: the outer block width (BHS_LU_NBO = 128 / 256 (default) / 384 / 512) only changes
    the blocking -- every variant solves to 1e-12 -- and applying the row interchanges once per outer block (default) or
    with every panel (BHS_LU_PERM_NOW=1) is the same arithmetic: bit-identical factors and solutions."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(N)
    A0 = torch.randn(S, N, N, dtype=torch.complex128, device="cuda", generator=g) + 0.1 * N ** 0.5 * torch.eye(N, dtype=torch.complex128, device="cuda")
    b0 = torch.randn(S, N, dtype=torch.complex128, device="cuda", generator=g)

    def solve():
        A, b = A0.clone(), b0.clone()
        ops.zgesv_batched_(A, b, ops.SolveBuffers(N, 1, S))
        torch.cuda.synchronize()
        return A, b

    A_ref, x_ref = solve()
    def scaled_residual(x):  # |A x - b| / (|A| |x|): the backward-error measure that does not depend on the conditioning
        r = torch.linalg.norm(torch.einsum("sij,sj->si", A0, x) - b0, dim=1)
        return (r / (torch.linalg.matrix_norm(A0) * torch.linalg.norm(x, dim=1))).max().item()

    assert scaled_residual(x_ref) < 1e-14
    monkeypatch.setenv("BHS_LU_PERM_NOW", "1")
    A_now, x_now = solve()
    monkeypatch.delenv("BHS_LU_PERM_NOW")
    assert torch.equal(A_now, A_ref) and torch.equal(x_now, x_ref)
    for nbo in (128, 384, 512):
        monkeypatch.setenv("BHS_LU_NBO", str(nbo))
        _, x = solve()
        monkeypatch.delenv("BHS_LU_NBO")
        err = (torch.linalg.norm(x - x_ref) / torch.linalg.norm(x_ref)).item()
        print(f"\nN={N} S={S} nbo={nbo}: scaled residual {scaled_residual(x):.2e}, rel. difference to the default blocking {err:.2e}")
        assert scaled_residual(x) < 1e-14 and err < 1e-9


def test_zgesv_needs_pivoting(ops):
    import torch

    N = 300
    g = torch.Generator(device="cuda").manual_seed(5)
    perm = torch.randperm(N, device="cuda", generator=g)
    A = torch.zeros(N, N, dtype=torch.complex128, device="cuda")
    A[torch.arange(N, device="cuda"), perm] = 1.0 + 0.5j
    A += 1e-3 * torch.randn(N, N, dtype=torch.complex128, device="cuda", generator=g)
    b = torch.randn(N, 3, dtype=torch.complex128, device="cuda", generator=g)
    x, bufs = ops.zgesv_(A.clone(), b.clone())
    want = torch.linalg.solve(A, b)
    assert int(bufs.info.item()) == 0
    assert float((x - want).abs().max() / want.abs().max()) < 1e-11


def test_zgetrf_zgetrs_and_singular(ops):
    import torch

    N = 160
    g = torch.Generator(device="cuda").manual_seed(6)
    A = torch.randn(N, N, dtype=torch.complex128, device="cuda", generator=g)
    b = torch.randn(N, 2, dtype=torch.complex128, device="cuda", generator=g)
    LU = A.clone()
    bufs = ops.zgetrf_(LU)
    x = ops.zgetrs_(LU, bufs, b.clone())
    assert float((A @ x - b).abs().max()) < 1e-9
    # more right-hand sides than one warp of the permutation kernel (every column must be permuted)
    b40 = torch.randn(N, 40, dtype=torch.complex128, device="cuda", generator=g)
    x40 = ops.zgetrs_(LU, bufs, b40.clone())
    assert float((A @ x40 - b40).abs().max()) < 1e-9
    x40s, _ = ops.zgesv_(A.clone(), b40.clone())
    assert float((A @ x40s - b40).abs().max()) < 1e-9
    S = torch.zeros(64, 64, dtype=torch.complex128, device="cuda")
    S[:40, :40] = torch.eye(40, dtype=torch.complex128, device="cuda")
    bufs2 = ops.zgetrf_(S)
    torch.cuda.synchronize()
    assert int(bufs2.info.item()) == 41


@pytest.mark.parametrize("d,n_end,B", [(2, 8, 3), (2, 40, 2), (3, 6, 2), (3, 16, 3), (3, 24, 2), (3, 30, 2), (4, 5, 2), (4, 10, 2)])
@pytest.mark.parametrize("mode", ["near", "per_ball", "far", "far_per_ball", "inner"])
def test_uscat(ops, d, n_end, B, mode):
    rng = np.random.default_rng(7)
    centers = np.zeros((B, d))
    centers[:, 1] = 4.0 * np.arange(B) - 2.0 * (B - 1)
    radii = rng.uniform(0.7, 1.1, size=B)
    k, eta = 1.3, 0.8
    H = bo.harm_count(d, n_end)
    deg = bo.degree_table(BT[d], n_end)
    density = (rng.normal(size=(B, H)) + 1j * rng.normal(size=(B, H))) * (0.5 ** deg)[None, :]
    P = 301
    x = rng.uniform(-6, 6, size=(d, P))
    x[:, 0] = 0.0
    x[:, 1] = centers[0] + np.eye(d)[0] * 2.0  # on a polar axis of ball 0
    if mode.startswith("far"):
        x = x / np.linalg.norm(x, axis=0, keepdims=True)
    if mode == "inner":
        x = centers[0][:, None] + rng.uniform(-0.5, 0.5, size=(d, P))
        centers, radii, density = centers[:1], radii[:1], density[:1]
    res = bo.OracleResult(c=bo.OracleCoordinates(BT[d]), centers=centers.T.copy(), radii=radii, k=k, n_end=n_end,
                          eta=eta, kind="inner" if mode == "inner" else "outer", density=density, matrix=None)
    far = mode.startswith("far")
    pb = mode.endswith("per_ball")
    want = bo.biem_u(res, x, far_field=far, per_ball=pb)
    got = ops.uscat(d, n_end, centers, radii, k, eta, density, x, far_field=far, per_ball=pb,
                    inner=(mode == "inner")).cpu().numpy()
    nan_w, nan_g = np.isnan(want), np.isnan(got)
    assert np.array_equal(nan_w, nan_g)
    if mode == "inner":
        # |h_n(kr)| explodes towards the centre; compare relative to the per-point magnitude
        ok = ~nan_w
        err = float(np.max(np.abs(got[ok] - want[ok]) / np.abs(want[ok])))
    else:
        err = relerr(got[~nan_w], want[~nan_w])
    print(f"\nuscat d={d} n_end={n_end} B={B} {mode}: rel err {err:.2e} (nan {int(nan_w.sum())})")
    assert err < 1e-11
