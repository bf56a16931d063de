"""One eager pass of the hot path for ncu: C3 system (rhs + assemble + LU solve + probe uscat) and a C5 field tile.

Run under `ncu --profile-from-start off`: the profiled range is bracketed with cudaProfilerStart/Stop after a warm-up.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops  # noqa: E402
from biem_helmholtz_sphere_b200.geometry import field_grid, grid_centers, probe_ring  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
n_end = 16
cen = torch.as_tensor(grid_centers(2, 3), device=dev)
B = cen.shape[0]
rad = torch.ones(B, dtype=torch.float64, device=dev)
H = n_end * n_end
N = B * H
k = torch.tensor([3.7], dtype=torch.float64, device=dev)
eta = torch.ones(1, dtype=torch.float64, device=dev)
dirv = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device=dev)
x = torch.as_tensor(probe_ring(64, 10.0, 3), device=dev)
A = torch.empty((1, N, N), dtype=torch.complex128, device=dev)
bufs = _ops.SolveBuffers(N, 1)

cen5 = torch.as_tensor(grid_centers(4, 3), device=dev)
rad5 = torch.ones(64, dtype=torch.float64, device=dev)
rng = np.random.default_rng(0)
deg = np.repeat(np.arange(24), 2 * np.arange(24) + 1)
dens5 = torch.as_tensor((rng.standard_normal((64, 576)) + 1j * rng.standard_normal((64, 576))) * np.exp(-0.7 * deg), device=dev)
x5 = torch.as_tensor(field_grid(grid, 20.0, 3).reshape(3, -1).copy(), device=dev)


def one():
    f = _ops.rhs_expand(3, n_end, centers=cen, radii=rad, k_in=k, direction=dirv)
    _ops.assemble(3, n_end, cen, rad, k, eta, out=A)
    r = f.reshape(N).clone()
    _ops.zgesv_(A[0], r, bufs)
    u = _ops.uscat(3, n_end, cen, rad, 3.7, 1.0, r.reshape(B, H), x)
    u5 = _ops.uscat(3, 24, cen5, rad5, 1.0, 1.0, dens5, x5)
    return u, u5


one()
torch.cuda.synchronize()
torch.cuda.profiler.start()
u, u5 = one()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", complex(u[0]), int(torch.isnan(u5.real).sum()))
