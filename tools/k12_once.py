"""A few launches of K1 (bhs_bessel) and K2 (bhs_harmonics) at the sizes bench.py times, for ncu and quick timing."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
nx, n_max = 1 << 21, 31
x = torch.rand(nx, dtype=torch.float64, device=dev, generator=g) * 40.0 + 0.5
npts, n_end = 1 << 18, 16
xyz = torch.randn(3, npts, dtype=torch.float64, device=dev, generator=g)
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = ev(lambda: _ops.bessel(3, 2, n_max, x))
print(f"K1 h_n n<=31, {nx} args: {ms*1e3:.0f} us, {nx*(n_max+1)*16/ms*1e-6:.0f} GB/s")
ms = ev(lambda: _ops.bessel(3, 0, n_max, x))
print(f"K1 j_n n<=31, {nx} args: {ms*1e3:.0f} us, {nx*(n_max+1)*16/ms*1e-6:.0f} GB/s")
ms = ev(lambda: _ops.harmonics(3, n_end, xyz))
print(f"K2 n_end 16, {npts} dirs: {ms*1e3:.0f} us, {npts*n_end*n_end*16/ms*1e-6:.0f} GB/s")
