"""cProfile of the host side of the C3 sweep through the public API with NumPy in / NumPy out (what bench.py's e2e times)."""
import cProfile, os, pstats, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biem_helmholtz_sphere_b200 as bhs
from biem_helmholtz_sphere_b200.geometry import grid_centers, probe_ring

K = 256
c = bhs.create_from_branching_types("ba")
cen = grid_centers(2, 3); rad = np.ones(cen.shape[0])
ks = np.linspace(0.5, 8.0, K); eta = np.ones(K)
d = np.array([[1.0], [0.0], [0.0]])
x = np.concatenate([np.zeros((3, 1)), probe_ring(64, 10.0, 3)], axis=1)
def step():
    uin, _ = bhs.plane_wave(k=ks, direction=d)
    res = bhs.biem(c, centers=cen[None], radii=rad[None], k=ks, n_end=16, eta=eta, uin=uin, keep_matrix=False)
    return res.density, res.uscat(x)
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print(f"one step: {(time.perf_counter()-t0)*1e3:.1f} ms")
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
