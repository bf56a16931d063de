"""Integer checksums of assembled matrices (bit-exact comparison of assembly kernels across processes:
run once plainly and once with BHS_ASM_LEGACY=1 and diff the output)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops
from biem_helmholtz_sphere_b200.geometry import grid_centers

dev = torch.device("cuda")
torch.manual_seed(0)
for d, half, n_end, nsys, jitter in ((3, 2, 16, 2, 0.0), (3, 1, 7, 3, 0.3), (2, 2, 20, 2, 0.0), (3, 2, 24, 1, 0.0), (4, 1, 4, 2, 0.1)):
    cen = torch.as_tensor(grid_centers(half, d), device=dev)
    B = cen.shape[0]
    if jitter:
        cen = cen + jitter * torch.rand_like(cen)
    rad = 0.5 + torch.rand(B, dtype=torch.float64, device=dev)
    k = torch.linspace(0.7, 2.9, nsys, dtype=torch.float64, device=dev)
    A = _ops.assemble(d, n_end, cen, rad, k, k)
    torch.cuda.synchronize()
    v = A.contiguous().view(torch.float64).view(torch.int64)
    print(d, B, n_end, nsys, int(v.sum().item()), int((v ^ (v >> 7)).sum().item()), bool(torch.isfinite(A.view(torch.float64)).all().item()))
