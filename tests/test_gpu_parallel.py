"""Sharded drivers of biem_helmholtz_sphere_b200.parallel on real GPUs: world = min(2, device_count) ranks over NCCL
(one process per GPU; a single process when the box has one GPU).  Default CUDA back ends, checked against the
oracle and against the unsharded public API."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


CEN = np.array([[0.0, 2.0, 0.0], [0.0, -2.0, 0.0], [4.0, 0.5, 1.0]])
RAD = np.array([1.0, 1.0, 0.7])
KS = np.linspace(0.6, 2.0, 5)
X = np.array([[0.0, 6.0, -3.0], [0.0, 0.5, 4.0], [0.0, 1.0, -2.0]])
N_END = 7


def _grid():
    g = np.linspace(-6.0, 6.0, 9)
    X0, X1 = np.meshgrid(g, g + 0.123, indexing="ij")
    return np.stack([X0, X1, np.full_like(X0, 0.3)])


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import biem_helmholtz_sphere_b200 as bhs
        from biem_helmholtz_sphere_b200 import parallel as par

        c = bhs.create_from_branching_types("ba")
        out = par.sweep(c, centers=CEN, radii=RAD, ks=KS, n_end=N_END, eta=1.0, x=X)
        dens1 = out["density"][1] if rank == 0 else None
        u = par.uscat_sharded(c, centers=CEN, radii=RAD, k=float(KS[1]), eta=1.0, n_end=N_END, density=dens1,
                              density_shape=(3, N_END * N_END), x_grid=_grid())
        # assembly of one system by block rows (bhs_assemble_rows) + all-gather of the strips: equals the unsharded matrix
        from biem_helmholtz_sphere_b200 import _ops

        A = par.assemble_sharded(c, centers=CEN, radii=RAD, k=1.7, eta=0.9, n_end=N_END, alpha=1.0, beta=0.2 + 0.1j)
        A1 = _ops.assemble(3, N_END, CEN, RAD, np.array([1.7]), np.array([0.9]), np.full(3, 1.0 + 0j), np.full(3, 0.2 + 0.1j))[0]
        assert A.shape == A1.shape and torch.equal(A.to(A1.device), A1)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), density=out["density"].cpu().numpy(),
                 uscat=out["uscat"].cpu().numpy(), field=u.cpu().numpy())
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_sweep_and_field_sharded(tmp_path):
    import torch
    import torch.multiprocessing as mp

    world = min(2, torch.cuda.device_count())
    if world > 1:
        mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    else:
        _worker(0, 1, _free_port(), str(tmp_path))
    from oracle import biem_oracle as O

    dens, us = [], []
    for k in KS:
        uin, _ = O.plane_wave(k=float(k), direction=np.array([1.0, 0.0, 0.0]))
        r = O.biem("ba", centers=CEN, radii=RAD, k=float(k), n_end=N_END, uin=uin, eta=1.0)
        dens.append(r.density)
        us.append(r.uscat(X))
        if len(dens) == 2:
            field = r.uscat(_grid())
    dens, us = np.stack(dens), np.stack(us)
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, f"r{rank}.npz"))
        assert np.max(np.abs(got["density"] - dens)) / np.max(np.abs(dens)) < 1e-10
        assert np.max(np.abs(got["uscat"] - us)) / np.max(np.abs(us)) < 1e-10
        nan = np.isnan(field)
        assert np.array_equal(nan, np.isnan(got["field"]))
        assert np.max(np.abs(got["field"][~nan] - field[~nan])) / np.max(np.abs(field[~nan])) < 1e-10
