"""A few eager assemblies of the C3 (or, with arguments `24 4`, C5) system, for ncu launch lists and quick timing."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from biem_helmholtz_sphere_b200 import _ops
from biem_helmholtz_sphere_b200.geometry import grid_centers

n_end = int(sys.argv[1]) if len(sys.argv) > 1 else 16
half = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
cen = torch.as_tensor(grid_centers(half, 3), device=dev)
B = cen.shape[0]
rad = torch.ones(B, dtype=torch.float64, device=dev)
k = torch.tensor([1.7], dtype=torch.float64, device=dev)
N = B * n_end * n_end
A = torch.empty((1, N, N), dtype=torch.complex128, device=dev)
work = _ops._work(_ops.load().bhs_assemble_workspace(_ops.get_plan(3, n_end).handle, B, 1))
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _ops.assemble(3, n_end, cen, rad, k, k, out=A, work=work)
    e1.record()
    torch.cuda.synchronize()
print(f"N = {N}: {e0.elapsed_time(e1)*1e3:.1f} us per bhs_assemble call (last of 3), {16.0*N*N/e0.elapsed_time(e1)*1e-6:.0f} GB/s")
