"""Time bhs_assemble alone at C3 (16 spheres, n_end 16) and C5 (64 spheres, n_end 24): best / median of 10 launches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops
from biem_helmholtz_sphere_b200.geometry import grid_centers

dev = torch.device("cuda")
for half, n_end in ((2, 16), (4, 24)):
    cen = torch.as_tensor(grid_centers(half, 3), device=dev)
    B = cen.shape[0]
    rad = torch.ones(B, dtype=torch.float64, device=dev)
    N = B * n_end * n_end
    k = torch.tensor([1.0], dtype=torch.float64, device=dev)
    A = torch.empty((1, N, N), dtype=torch.complex128, device=dev)
    work = _ops._work(_ops.load().bhs_assemble_workspace(_ops.get_plan(3, n_end).handle, B, 1))
    ts = []
    for it in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _ops.assemble(3, n_end, cen, rad, k, k, out=A, work=work)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    print(f"B={B} n_end={n_end} N={N}: best {ts[0]:.3f} ms, median {ts[len(ts)//2]:.3f} ms -> {16.0*N*N/ts[0]*1e-6:.0f} GB/s (whole call incl. pre-kernels)")
    del A
