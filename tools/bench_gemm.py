"""Trailing-update GEMM (bhs_zgemm_sub) on the shapes of a C3 factorisation, launched back to back; TFLOP/s = 8 M N K / t."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from biem_helmholtz_sphere_b200 import _ops

dev = torch.device("cuda")
C128 = torch.complex128
peak = _ops.fp64_peak(1, 4096)
print(f"DMMA peak {peak:.2f} TFLOP/s")
g = torch.Generator(device=dev).manual_seed(0)
for (M, N, K, reps) in ((3968, 3968, 128, 20), (2048, 2048, 128, 40), (3968, 96, 32, 50), (3968, 64, 64, 50), (64, 3968, 64, 50), (8064, 8064, 128, 5), (3968, 3968, 1024, 5)):
    A = torch.randn(M, K, dtype=C128, device=dev, generator=g)
    B = torch.randn(K, N, dtype=C128, device=dev, generator=g)
    Cm = torch.randn(M, N, dtype=C128, device=dev, generator=g)
    want = Cm - A @ B
    wk = _ops._work(_ops.load().bhs_zgemm_workspace(M, N, K))
    got = _ops.zgemm_sub_(Cm.clone(), A, B, work=wk)
    err = float((got - want).abs().max() / want.abs().max())
    for _ in range(3):
        _ops.zgemm_sub_(Cm, A, B, work=wk)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _ops.zgemm_sub_(Cm, A, B, work=wk)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 8.0 * M * N * K / ms * 1e-9
    print(f"{M}x{N}x{K}: {ms*1e3:8.1f} us  {tf:6.2f} TFLOP/s  ({tf/peak:.3f} of peak)  err {err:.1e}", flush=True)
    del A, B, Cm, want, got
