import torch, time, os, sys
sys.path.insert(0, "/root/repo")
import biem_helmholtz_sphere_b200 as bhs
from biem_helmholtz_sphere_b200.geometry import grid_centers, sweep_wavenumbers
dev = torch.device("cuda")
c = bhs.create_from_branching_types("ba")
cen = torch.as_tensor(grid_centers(2,3), device=dev); rad = torch.ones(16, dtype=torch.float64, device=dev)
ks = torch.as_tensor(sweep_wavenumbers(256), device=dev); eta = torch.ones(256, dtype=torch.float64, device=dev)
d = torch.tensor([[1.0],[0.0],[0.0]], dtype=torch.float64, device=dev)
def step():
    uin,_ = bhs.plane_wave(k=ks, direction=d)
    return bhs.biem(c, centers=cen[None], radii=rad[None], k=ks, n_end=16, eta=eta, uin=uin, keep_matrix=False)
for _ in range(3): step()
torch.cuda.synchronize(); t=time.time(); step(); step(); torch.cuda.synchronize(); dt=(time.time()-t)/2
print(os.environ.get("BHS_LU_EXPERIMENT"), os.environ.get("BHS_LU_GEMM_ONLY"), "systems/s", 256/dt)
