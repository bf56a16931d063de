"""Single-system LU latency (bhs_zgesv, one right-hand side) at a few sizes: eager launches and a CUDA-graph replay."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from biem_helmholtz_sphere_b200 import _ops

dev = torch.device("cuda")
for N in (1024, 4096, 8192):
    g = torch.Generator(device=dev).manual_seed(N)
    A0 = torch.randn(N, N, dtype=torch.complex128, device=dev, generator=g)
    b0 = torch.randn(N, 1, dtype=torch.complex128, device=dev, generator=g)
    A, b = A0.clone(), b0.clone()
    bufs = _ops.SolveBuffers(N, 1)
    _ops.zgesv_(A, b, bufs)
    res = (torch.linalg.norm(A0 @ b - b0) / torch.linalg.norm(b0)).item()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 5
    tot = 0.0
    for _ in range(reps):
        A.copy_(A0); b.copy_(b0)
        ev[0].record(); _ops.zgesv_(A, b, bufs); ev[1].record(); torch.cuda.synchronize()
        tot += ev[0].elapsed_time(ev[1])
    eager = tot / reps
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        A.copy_(A0); b.copy_(b0)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            _ops.zgesv_(A, b, bufs)
        tot = 0.0
        for _ in range(reps):
            A.copy_(A0); b.copy_(b0)
            ev[0].record(s); gr.replay(); ev[1].record(s); s.synchronize()
            tot += ev[0].elapsed_time(ev[1])
    graph = tot / reps
    fl = 8.0 / 3.0 * N ** 3
    print(f"N={N}: residual {res:.2e}  eager {eager:.2f} ms ({fl/eager*1e-9:.2f} TF)  graph {graph:.2f} ms ({fl/graph*1e-9:.2f} TF)")
