"""One eager pass of what a sweep group runs -- S systems (default 4) assembled by one bhs_assemble call and factorised in
lock step by bhs_zgesv_batched -- for ncu launch lists / metric captures of the sweep's kernels.

Run under `ncu --profile-from-start off`: the profiled range is bracketed with cudaProfilerStart/Stop after a warm-up pass."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops  # noqa: E402
from biem_helmholtz_sphere_b200.geometry import grid_centers  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
n_end = 16
cen = torch.as_tensor(grid_centers(2, 3), device=dev)
B = cen.shape[0]
rad = torch.ones(B, dtype=torch.float64, device=dev)
N = B * n_end * n_end
ks = torch.linspace(0.5, 8.0, 256, dtype=torch.float64, device=dev)[:S].contiguous()
eta = torch.ones(S, dtype=torch.float64, device=dev)
dirv = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device=dev)
A = torch.empty((S, N, N), dtype=torch.complex128, device=dev)
bufs = _ops.SolveBuffers(N, 1, S)
work = _ops._work(_ops.load().bhs_assemble_workspace(_ops.get_plan(3, n_end).handle, B, S))


def one():
    f = _ops.rhs_expand(3, n_end, centers=cen, radii=rad, k_in=ks, direction=dirv)
    _ops.assemble(3, n_end, cen, rad, ks, eta, out=A, work=work)
    r = f.reshape(S, N).clone()
    _ops.zgesv_batched_(A, r, bufs)
    return r


one()
torch.cuda.synchronize()
torch.cuda.profiler.start()
r = one()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", bool(torch.isfinite(r.view(torch.float64)).all()))
