"""One trailing-update GEMM launch for ncu (3968 x 3968 x K, default K = 128), after a warm-up launch."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from biem_helmholtz_sphere_b200 import _ops
K = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = int(sys.argv[2]) if len(sys.argv) > 2 else 3968
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(M, K, dtype=torch.complex128, device=dev, generator=g)
B = torch.randn(K, M, dtype=torch.complex128, device=dev, generator=g)
Cm = torch.randn(M, M, dtype=torch.complex128, device=dev, generator=g)
wk = _ops._work(_ops.load().bhs_zgemm_workspace(M, M, K))
_ops.zgemm_sub_(Cm, A, B, work=wk)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
_ops.zgemm_sub_(Cm, A, B, work=wk)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
