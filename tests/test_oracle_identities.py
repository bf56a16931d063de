"""First-principles checks of the oracle that do not go through the scalar uscat(0) of the golden CSVs: the harmonics
are orthonormal under the oracle's own quadrature, and the (S|R) translation matrix satisfies the addition theorem
    S_{h'}(z + t) = sum_h (S|R)_{h',h}(t) R_h(z),   |z| < |t|          (_biem.py:522)
point-wise.  These pin the matrix entries / index conventions that the golden values (basis-invariant scalars) cannot."""
import numpy as np
import pytest

from oracle import biem_oracle as bo


@pytest.mark.parametrize("btype,n_end", [("a", 9), ("ba", 7), ("bba", 5), ("bbba", 4), ("bbbba", 3)])
def test_harmonics_orthonormal(btype, n_end):
    angles, w = bo.quadrature(btype, 2 * n_end)  # exact for products of two harmonics of degree < n_end
    Y = bo.harmonics(btype, angles, n_end)  # [Q, H]
    G = (Y.conj().T * w) @ Y
    assert np.allclose(G, np.eye(G.shape[0]), atol=1e-12)


@pytest.mark.parametrize("btype,n_end,k", [("a", 14, 1.3), ("ba", 12, 0.9), ("ba", 12, 2.0 + 0.5j), ("bba", 8, 1.1),
                                           ("bbba", 6, 1.2), ("bbbba", 5, 0.8)])
def test_translation_addition_theorem(btype, n_end, k):
    d = len(btype) + 1
    rng = np.random.default_rng(5)
    coords = bo.OracleCoordinates(btype)
    t = rng.normal(size=d)
    t *= 4.0 / np.linalg.norm(t)
    T = bo.translation_coef(btype, t[:, None], k, n_end)[0]  # [H', H]
    deg = bo.degree_table(btype, n_end)
    nlow = 4 if d < 5 else 2  # singular functions S_{h'} of low degree: the series in h converges like (|z| / |t|)^n
    for _ in range(5):
        z = rng.normal(size=d)
        z *= (0.35 if d < 4 else 0.1 if d == 4 else 0.03) / np.linalg.norm(z)  # truncation ~ (|z| / |t|)^n_end: shorter series in high d
        sz = coords.from_cartesian((z + t)[:, None])
        S = (bo.radial(d, n_end - 1, k * sz["r"], "h")[deg, 0] * bo.harmonics(btype, [sz[i] for i in range(d - 1)], n_end)[0])
        rz = coords.from_cartesian(z[:, None])
        R = (bo.radial(d, n_end - 1, k * rz["r"], "j")[deg, 0] * bo.harmonics(btype, [rz[i] for i in range(d - 1)], n_end)[0])
        lhs = S[deg < nlow]
        rhs = (T @ R)[deg < nlow]
        assert np.max(np.abs(lhs - rhs)) < 1e-9 * np.max(np.abs(lhs)), (btype, np.max(np.abs(lhs - rhs)))


@pytest.mark.parametrize("btype,n_end,k", [("a", 12, 1.7), ("ba", 9, 0.9), ("ba", 9, 1.4 + 0.3j), ("bba", 6, 1.1), ("bbba", 4, 1.2)])
def test_translation_block_sign_symmetry(btype, n_end, k):
    """(S|R)_{h',h}(-t) = (-1)^(n + n') (S|R)_{h',h}(t): every term of an entry has n'' = n + n' (mod 2) and
    S_{h''}(-t) = (-1)^{n''} S_{h''}(t).  The assembly kernel relies on it: the block of the ball pair (b', b) is derived from
    the block of (b, b') by that sign pattern (csrc/assemble.cu, pair_rep_kernel / assemble_reg_kernel), for any geometry."""
    d = len(btype) + 1
    rng = np.random.default_rng(11)
    deg = bo.degree_table(btype, n_end)
    sign = np.where((deg[:, None] + deg[None, :]) % 2 == 0, 1.0, -1.0)
    for _ in range(3):
        t = rng.normal(size=d)
        t *= rng.uniform(2.5, 6.0) / np.linalg.norm(t)
        Tp = bo.translation_coef(btype, t[:, None], k, n_end)[0]
        Tm = bo.translation_coef(btype, -t[:, None], k, n_end)[0]
        scale = np.max(np.abs(Tp))
        assert np.max(np.abs(Tm - sign * Tp)) < 1e-12 * scale
