"""Determinism / race stress of the DMMA kernel: the same C -= A B on several streams at once, many times; every
result must be bit-identical to the first and close to a torch reference."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops

torch.manual_seed(0)
dev = torch.device("cuda")
M, N, K = 3968, 3968, 128
A = torch.randn(M, K, dtype=torch.complex128, device=dev)
B = torch.randn(K, N, dtype=torch.complex128, device=dev)
C0 = torch.randn(M, N, dtype=torch.complex128, device=dev)
ref = C0 - A @ B
streams = [torch.cuda.Stream() for _ in range(6)]
works = [_ops._work(_ops.load().bhs_zgemm_workspace(M, N, K)) for _ in streams]
outs = []
torch.cuda.synchronize()
for rep in range(5):
    for st, w in zip(streams, works):
        with torch.cuda.stream(st):
            C = C0.clone()
            _ops.zgemm_sub_(C, A, B, work=w)
            outs.append(C)
torch.cuda.synchronize()
err = float((outs[0] - ref).abs().max() / ref.abs().max())
nbad = sum(int(not torch.equal(o, outs[0])) for o in outs)
worst = max(float((o - outs[0]).abs().max()) for o in outs)
print(f"rel err vs torch {err:.2e}; {nbad} of {len(outs)} results differ from the first (max abs diff {worst:.2e})")
sys.exit(1 if (nbad or err > 1e-13) else 0)
