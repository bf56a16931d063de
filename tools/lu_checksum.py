"""Integer checksum of the LU factors and the solution of a seeded random system through bhs_zgesv (lone-system path), to
compare library variants bit for bit across processes (e.g. BHS_LU_FUSE_PERM=0 vs default)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops

for N in (700, 4096, 6000):
    g = torch.Generator(device="cuda").manual_seed(N)
    A = torch.randn(N, N, dtype=torch.complex128, device="cuda", generator=g)
    b = torch.randn(N, 1, dtype=torch.complex128, device="cuda", generator=g)
    A0, b0 = A.clone(), b.clone()
    _ops.zgesv_(A, b, _ops.SolveBuffers(N, 1))
    torch.cuda.synchronize()
    res = (torch.linalg.norm(A0 @ b - b0) / torch.linalg.norm(b0)).item()
    ca = int(A.view(torch.float64).view(torch.int64).sum().item())
    cb = int(b.view(torch.float64).view(torch.int64).sum().item())
    print(N, ca, cb, f"{res:.2e}")
