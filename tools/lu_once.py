"""`reps` single-system solves of size N (for ncu launch lists: python tools/lu_once.py 4096 2)."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from biem_helmholtz_sphere_b200 import _ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(N)
A0 = torch.randn(N, N, dtype=torch.complex128, device=dev, generator=g)
b0 = torch.randn(N, 1, dtype=torch.complex128, device=dev, generator=g)
bufs = _ops.SolveBuffers(N, 1)
for _ in range(reps):
    A, b = A0.clone(), b0.clone()
    _ops.launch_count(reset=True)
    _ops.zgesv_(A, b, bufs)
    torch.cuda.synchronize()
    print("launches", _ops.launch_count())
