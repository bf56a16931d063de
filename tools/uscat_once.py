"""One planar and one general launch of the 3-D field kernels on a 1024 x 1024 tile of the C5 grid (64 spheres, n_end = 24),
for ncu:  ncu --set full --profile-from-start off -k regex:uscat3d python tools/uscat_once.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, "/root/repo")
from biem_helmholtz_sphere_b200 import _ops
from biem_helmholtz_sphere_b200.geometry import field_grid, grid_centers

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_end, half = 24, 4
dev = torch.device("cuda")
cen = torch.as_tensor(grid_centers(half, 3), device=dev)
B = cen.shape[0]
rad = torch.ones(B, dtype=torch.float64, device=dev)
H = n_end * n_end
rng = np.random.default_rng(0)
deg = np.repeat(np.arange(n_end), 2 * np.arange(n_end) + 1)
dens = torch.as_tensor((rng.standard_normal((B, H)) + 1j * rng.standard_normal((B, H))) * np.exp(-0.7 * deg)[None, :], device=dev)
x = torch.as_tensor(field_grid(G, 20.0, 3).reshape(3, -1), device=dev).contiguous()
xg = x.clone()
xg[2] += 0.37
work = _ops._work(_ops.load().bhs_uscat_workspace(_ops.get_plan(3, n_end).handle, B))
for pts in (x, xg):
    _ops.uscat(3, n_end, cen, rad, 1.0, 1.0, dens, pts, work=work)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for pts in (x, xg):
    _ops.uscat(3, n_end, cen, rad, 1.0, 1.0, dens, pts, work=work)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
