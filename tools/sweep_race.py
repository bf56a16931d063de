"""Run the same C3 sweep several times through biem() and compare the densities bit for bit (the sweep engine's streams,
graphs and batched kernels are deterministic: any difference is a race).  python tools/sweep_race.py [K] [reps]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biem_helmholtz_sphere_b200 as bhs
from biem_helmholtz_sphere_b200.geometry import grid_centers

K = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda")
c = bhs.create_from_branching_types("ba")
cen = torch.as_tensor(grid_centers(2, 3), device=dev)
rad = torch.ones(cen.shape[0], dtype=torch.float64, device=dev)
ks = torch.linspace(0.5, 8.0, K, dtype=torch.float64, device=dev)
d = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device=dev)
ref = None
for r in range(reps):
    uin, _ = bhs.plane_wave(k=ks, direction=d[:, None])
    res = bhs.biem(c, centers=cen[None], radii=rad[None], k=ks, n_end=16, eta=torch.ones_like(ks), uin=uin, keep_matrix=False)
    dens = res.density.reshape(K, -1).clone()
    torch.cuda.synchronize()
    if ref is None:
        ref = dens
        print(f"rep 0: reference, finite {bool(torch.isfinite(dens.view(torch.float64)).all())}")
        continue
    same = torch.equal(dens.view(torch.float64).view(torch.int64), ref.view(torch.float64).view(torch.int64))
    if same:
        print(f"rep {r}: bit-identical")
    else:
        rel = ((dens - ref).abs().amax(1) / ref.abs().amax(1)).cpu().numpy()
        bad = np.nonzero(rel > 0)[0]
        print(f"rep {r}: {bad.size} systems differ, max rel {rel.max():.2e}, systems {bad[:16].tolist()}")
