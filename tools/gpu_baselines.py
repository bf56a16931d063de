"""Library reference points on the box (design aid + numbers quoted in DESIGN.md): the reference's own GPU path is
torch.linalg.solve -> cuSOLVER (_biem.py:797 with torch CUDA tensors), and cuBLAS ZGEMM is the cross-check of the FP64
tensor peak that bhs_fp64_peak measures."""
import json
import sys
import time

import torch

sys.path.insert(0, "/root/repo")
dev = torch.device("cuda")
C128 = torch.complex128
out = {}


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device=dev).manual_seed(0)
for N in (4096, 8192):
    A = torch.randn(N, N, dtype=C128, device=dev, generator=g)
    b = torch.randn(N, 1, dtype=C128, device=dev, generator=g)
    ms = ev_time(lambda: torch.linalg.solve(A, b))
    ms_lu = ev_time(lambda: torch.linalg.lu_factor(A))
    fl = 8.0 / 3.0 * N**3
    out[f"torch_linalg_solve_N{N}"] = {"ms": ms, "tflops": fl / ms * 1e-9}
    out[f"torch_lu_factor_N{N}"] = {"ms": ms_lu, "tflops": fl / ms_lu * 1e-9}
    print(N, out[f"torch_linalg_solve_N{N}"], out[f"torch_lu_factor_N{N}"], flush=True)
# batched (the reference with a batched k would issue this)
N = 4096
for S in (8, 32):
    A = torch.randn(S, N, N, dtype=C128, device=dev, generator=g)
    b = torch.randn(S, N, 1, dtype=C128, device=dev, generator=g)
    ms = ev_time(lambda: torch.linalg.solve(A, b), reps=2, warm=1)
    out[f"torch_linalg_solve_batched_{S}xN{N}"] = {"ms": ms, "systems_per_s": S / ms * 1e3, "tflops": S * 8.0 / 3.0 * N**3 / ms * 1e-9}
    print(S, out[f"torch_linalg_solve_batched_{S}xN{N}"], flush=True)
    del A, b
# multi-stream loop of single solves (closest library analogue of our sweep engine)
streams = [torch.cuda.Stream() for _ in range(8)]
As = [torch.randn(N, N, dtype=C128, device=dev, generator=g) for _ in range(8)]
bs = [torch.randn(N, 1, dtype=C128, device=dev, generator=g) for _ in range(8)]
def multi():
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            torch.linalg.solve(As[i], bs[i])
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
ms = ev_time(multi, reps=3, warm=1)
out["torch_linalg_solve_8streams_N4096"] = {"ms": ms, "systems_per_s": 8 / ms * 1e3}
print(out["torch_linalg_solve_8streams_N4096"], flush=True)
del As, bs
# cuBLAS ZGEMM
for (M, Nn, K) in ((8192, 8192, 8192), (3968, 3968, 128), (4096, 4096, 256)):
    a = torch.randn(M, K, dtype=C128, device=dev, generator=g)
    bb = torch.randn(K, Nn, dtype=C128, device=dev, generator=g)
    c = torch.randn(M, Nn, dtype=C128, device=dev, generator=g)
    ms = ev_time(lambda: torch.addmm(c, a, bb, alpha=-1.0, out=c), reps=5 if M < 8192 else 2, warm=1)
    out[f"cublas_zgemm_{M}x{Nn}x{K}"] = {"ms": ms, "tflops": 8.0 * M * Nn * K / ms * 1e-9}
    print(out[f"cublas_zgemm_{M}x{Nn}x{K}"], flush=True)
    del a, bb, c
# real DGEMM
a = torch.randn(8192, 8192, dtype=torch.float64, device=dev); bb = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
ms = ev_time(lambda: a @ bb, reps=3, warm=1)
out["cublas_dgemm_8192"] = {"ms": ms, "tflops": 2.0 * 8192**3 / ms * 1e-9}
print(out["cublas_dgemm_8192"], flush=True)
del a, bb
from biem_helmholtz_sphere_b200 import _ops
out["bhs_fp64_peak"] = {"dfma": _ops.fp64_peak(0, 4096), "dmma884": _ops.fp64_peak(1, 4096)}
print(out["bhs_fp64_peak"], flush=True)
json.dump(out, open("/root/repo/gpurun_out/gpu_baselines.json", "w"), indent=1)
