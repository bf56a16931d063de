"""Host-side sharding logic (biem_helmholtz_sphere_b200.parallel) on CPU: pure index math, and world_size-2 gloo runs in
which the oracle is injected as the per-shard solver / tile evaluator (the CUDA path is the default back end in
production; here only the sharding, broadcast and gather are under test)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from biem_helmholtz_sphere_b200 import parallel as par  # noqa: E402


def test_shard_indices_partition():
    for K in (0, 1, 5, 32, 256, 257):
        for world in (1, 2, 3, 4, 8):
            parts = [par.shard_indices(K, r, world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts))
            assert np.array_equal(allidx, np.arange(K))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_indices(4, 2, 2)


def test_field_rows_partition():
    for n in (0, 1, 7, 2048):
        for world in (1, 2, 3, 8):
            sl = [par.field_rows(n, r, world) for r in range(world)]
            assert sl[0].start == 0 and sl[-1].stop == n
            for a, b in zip(sl[:-1], sl[1:]):
                assert a.stop == b.start
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_solve_shard(c, centers, radii, ks, n_end, eta, direction, x):
    from oracle import biem_oracle as O

    dens, us = [], []
    for k in ks:
        uin, _ = O.plane_wave(k=float(k), direction=np.asarray(direction))
        r = O.biem(c, centers=np.asarray(centers), radii=np.asarray(radii), k=float(k), n_end=n_end, uin=uin, eta=eta)
        dens.append(r.density)
        if x is not None:
            us.append(r.uscat(np.asarray(x)))
    return torch.as_tensor(np.stack(dens)), (torch.as_tensor(np.stack(us)) if x is not None else None)


def _oracle_eval_tile(c, centers, radii, k, eta, n_end, density, x_tile):
    from oracle import biem_oracle as O

    res = O.OracleResult(c=O.OracleCoordinates(c), centers=np.asarray(centers).T.copy(), radii=np.asarray(radii), k=float(k),
                         n_end=n_end, eta=float(eta), kind="outer", density=density.numpy(), matrix=None)
    return torch.as_tensor(res.uscat(np.asarray(x_tile)))


def _oracle_assemble_rows(c, centers, radii, k, eta, n_end, alpha, beta, b_lo, b_hi):
    from oracle import biem_oracle as O

    B = len(radii)
    A = O.assemble(c, np.asarray(centers), np.asarray(radii), k, n_end, eta, np.broadcast_to(np.asarray(alpha, complex), (B,)),
                   np.broadcast_to(np.asarray(beta, complex), (B,)))  # [B, H, B, H]
    H = A.shape[1]
    return torch.as_tensor(A.reshape(B * H, B * H)[b_lo * H : b_hi * H].copy())


def _worker(rank, world, port, K, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cen = np.array([[0.0, 2.0], [0.0, -2.0], [4.0, 0.5]])
        rad = np.array([1.0, 1.0, 0.7])
        ks = np.linspace(0.6, 2.0, K)
        x = np.array([[0.0, 6.0, -3.0], [0.0, 0.5, 4.0]])
        # (1) k-sweep: shard, solve locally, all-gather
        out = par.sweep("a", centers=cen, radii=rad, ks=ks, n_end=5, eta=1.0, x=x, solve_shard=_oracle_solve_shard)
        # (2) field evaluation: density only on rank 0, broadcast, row tiles, gather
        g = np.linspace(-6.0, 6.0, 7)
        X0, X1 = np.meshgrid(g, g + 0.123, indexing="ij")
        grid = np.stack([X0, X1])
        j = min(1, K - 1)
        dens0 = out["density"][j] if rank == 0 else None
        u = par.uscat_sharded("a", centers=cen, radii=rad, k=float(ks[j]), eta=1.0, n_end=5, density=dens0,
                              density_shape=tuple(out["density"][j].shape), x_grid=grid, eval_tile=_oracle_eval_tile)
        # (2b) assembly of one system by block rows + all-gather of the strips
        A = par.assemble_sharded("a", centers=cen, radii=rad, k=1.4, eta=1.0, n_end=5, alpha=1.0, beta=0.3,
                                 assemble_rows=_oracle_assemble_rows)
        A_ref = _oracle_assemble_rows("a", cen, rad, 1.4, 1.0, 5, 1.0, 0.3, 0, 3)
        assert A.shape == A_ref.shape and torch.equal(A, A_ref)
        strip = par.assemble_sharded("a", centers=cen, radii=rad, k=1.4, eta=1.0, n_end=5, gather=False,
                                     assemble_rows=_oracle_assemble_rows)
        lo, hi = par.ball_rows(3, rank, world)
        assert strip.shape[0] == (hi - lo) * 9
        # (3) local-only mode
        loc = par.sweep("a", centers=cen, radii=rad, ks=ks, n_end=5, eta=1.0, x=None, gather=False,
                        solve_shard=_oracle_solve_shard)
        assert np.array_equal(loc["indices"], par.shard_indices(K, rank, world))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), density=out["density"].numpy(), uscat=out["uscat"].numpy(),
                 field=u.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [5, 1])
def test_sweep_and_field_sharding_gloo_world2(tmp_path, K):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, K, str(tmp_path)), nprocs=world, join=True)
    # serial reference with the same injected solver
    cen = np.array([[0.0, 2.0], [0.0, -2.0], [4.0, 0.5]])
    rad = np.array([1.0, 1.0, 0.7])
    ks = np.linspace(0.6, 2.0, K)
    x = np.array([[0.0, 6.0, -3.0], [0.0, 0.5, 4.0]])
    dens, us = _oracle_solve_shard("a", cen, rad, ks, 5, 1.0, np.array([1.0, 0.0]), x)
    r0 = np.load(os.path.join(tmp_path, "r0.npz"))
    r1 = np.load(os.path.join(tmp_path, "r1.npz"))
    for r in (r0, r1):
        assert np.array_equal(r["density"], dens.numpy())  # gather must be bit-exact
        assert np.array_equal(r["uscat"], us.numpy())
    assert np.array_equal(r0["field"], r1["field"], equal_nan=True)
    j = min(1, K - 1)
    g = np.linspace(-6.0, 6.0, 7)
    X0, X1 = np.meshgrid(g, g + 0.123, indexing="ij")
    ref = _oracle_eval_tile("a", cen, rad, float(ks[j]), 1.0, 5, dens[j], np.stack([X0, X1]))
    if True:
        assert np.array_equal(r0["field"], ref.numpy().reshape(7, 7), equal_nan=True)
